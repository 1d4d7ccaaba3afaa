#!/usr/bin/env python
"""Null kernel time against the number of permutations (config B's size, long-tailed types, observed kernels off):
how the persistent grid of 296 CTAs fills its rounds.   python profiles/time_null_vs_perms.py [P ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth
from pctsea_parent_b200._lib import Context
n_used, T = 219530, 100
perms = [int(a) for a in sys.argv[1:]] or [148, 296, 444, 592, 888, 1000, 1184, 1480, 2000, 2960]
s, lab = synth.synthetic_ranked(n_used + 1, T, seed=1)
order = np.arange(n_used + 1, dtype=np.int32)
ctx = Context(0)
ctx.set_option(P.OPT_OBS_OVERLAP, 0)
for Pn in perms:
    prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=3)
    ts = []
    for _ in range(6):
        ctx.enrich(s, order, lab, T, prm, want_null=False)
        ts.append(ctx.timings()["null_ms"])
    t = min(ts[2:])
    print(f"P={Pn:5d}  rounds of 296 = {Pn / 296:5.2f}  null {t:.3f} ms  = {1e3 * t / Pn:.3f} us per permutation", flush=True)
