#!/usr/bin/env python
"""a10-a12 alone on a synthetic null [T][P] (run under `ncu --metrics gpu__time_duration.sum` for per-kernel times):
    python profiles/time_reduce.py T P [P ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pctsea_parent_b200 as P
from pctsea_parent_b200._lib import TYPE_RESULT_DTYPE
T = int(sys.argv[1])
ctx = P.Context(0)
rng = np.random.default_rng(0)
for Pn in [int(a) for a in sys.argv[2:]]:
    null = (rng.standard_normal((T, Pn)) * 0.1).astype(np.float32)
    res = np.zeros(T, TYPE_RESULT_DTYPE)
    res["ews"] = (rng.standard_normal(T) * 0.2).astype(np.float32)
    for _ in range(3):
        ctx.reduce_null(res, null)
        t = ctx.timings()
    print(f"T={T} P={Pn}: reduce {t['reduce_ms']:.3f} ms  total {t['total_ms']:.3f} ms", flush=True)
