#!/usr/bin/env python
"""Time the label + null KS stages alone on a synthetic ranked list: python profiles/time_null.py N_USED T P"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth
from pctsea_parent_b200._lib import Context
n_used, T, Pn = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
s, lab = synth.synthetic_ranked(n_used + 1, T, seed=1)
order = np.arange(n_used + 1, dtype=np.int32)
ctx = Context(0)
prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=3)
out = []
for _ in range(4):
    ctx.enrich(s, order, lab, T, prm, want_null=False)
    t = ctx.timings()
    out.append((t["labels_ms"], t["null_ms"], t["observed_ms"]))
print(f"Nu={n_used} T={T} P={Pn}: labels {out[-1][0]:.3f} ms  null {out[-1][1]:.3f} ms  observed {out[-1][2]:.3f} ms")
