#!/usr/bin/env python
"""Time the null stage alone on a synthetic ranked list, over kernel geometries:
    python profiles/time_null.py N_USED T P [threads,ctas_per_sm,min_threads,overlap ...]
Each geometry argument sets the ctx options OPT_NULL_THREADS, OPT_NULL_CTAS_PER_SM, OPT_NULL_MIN_THREADS,
OPT_OBS_OVERLAP (0 = library default for the first three). LONGTAIL=1: long-tailed cell-type sizes like a real ranking's."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth
from pctsea_parent_b200._lib import Context
n_used, T, Pn = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
geoms = [tuple(int(v) for v in a.split(",")) for a in sys.argv[4:]] or [(0, 0, 0, 1)]
s, lab = synth.synthetic_ranked(n_used + 1, T, seed=1, simple=bool(int(os.environ.get("SIMPLE", "0"))))
if int(os.environ.get("LONGTAIL", "0")):
    # cell-type sizes as long-tailed as a real ranking's (config B: 100 types, min 28, median 87, max 113 640 of 219 530):
    # sizes ~ 1 / k^2 with a floor of 28 cells, 0.1 % unknown
    rng = np.random.default_rng(7)
    w = 1.0 / np.arange(1, T + 1) ** 2
    sizes = np.maximum(28, np.floor(w / w.sum() * (n_used + 1 - 28 * T)).astype(np.int64))
    lab = np.repeat(np.arange(T, dtype=np.int32), sizes)[: n_used + 1]
    lab = np.concatenate([lab, np.full(n_used + 1 - lab.size, 0, np.int32)])
    rng.shuffle(lab)
    lab[rng.random(n_used + 1) < 0.001] = -1
order = np.arange(n_used + 1, dtype=np.int32)
ctx = Context(0)
prm = ctx.null_params(Pn, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=3)
for g in geoms:
    thr, ctas, minthr, ov = (list(g) + [0, 0, 0, 1])[:4] if len(g) < 4 else g
    ctx.set_option(P.OPT_NULL_THREADS, thr)
    ctx.set_option(P.OPT_NULL_CTAS_PER_SM, ctas)
    ctx.set_option(P.OPT_NULL_MIN_THREADS, minthr)
    ctx.set_option(P.OPT_OBS_OVERLAP, ov)
    out = []
    for _ in range(4):
        ctx.enrich(s, order, lab, T, prm, want_null=False)
        t = ctx.timings()
        out.append((t["null_ms"], t["observed_ms"], t["total_ms"]))
    print(f"Nu={n_used} T={T} P={Pn} threads={thr} ctas/SM={ctas} min_threads={minthr} overlap={ov}: "
          f"null {out[-1][0]:.3f} ms  observed {out[-1][1]:.3f} ms  total {out[-1][2]:.3f} ms", flush=True)
