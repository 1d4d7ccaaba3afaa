#!/usr/bin/env python
"""Which hash may the Feistel round function of the label permutation use? Null of the enrichment score (oracle, CPU,
EXACT arithmetic) under the keyed permutation of the size-sorted arrangement with a candidate hash, against the JDK
Collections.shuffle null: per-type log ratio of the standard deviations (the statistic that exposed 3-4 rounds: +2.7 %),
two-sample KS p-values, tail counts above the pooled 99th percentile.
    python profiles/feistel_mix_check.py N_USED T P VARIANTS ROUNDS      e.g.  1500 7 40000 0,1,3 5
Variants (oracle hook orc_set_mix_variant): 0 = production, t = v * M + key; t * (t >> 16) (three GPU instructions);
1 = multiply - xorshift - multiply (four; round 2's first form); 3 = t * t (two).
Output of the committed runs: profiles/feistel_mix_check.txt"""
import os, sys, ctypes as C, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import ks_2samp
from oracle import oracle as O
from pctsea_parent_b200 import synth
O.build()
L = O.lib()
Nu = int(sys.argv[1]); T = int(sys.argv[2]); P = int(sys.argv[3])
variants = [int(v) for v in sys.argv[4].split(',')]
rounds_list = [int(v) for v in sys.argv[5].split(',')]
s, lab = synth.synthetic_ranked(Nu, T, seed=21)
th = os.cpu_count()
t0=time.time()
jdk, _ = O.null(s, lab, T, P, perm_mode=0, seed=99, nthreads=th)
jdk2, _ = O.null(s, lab, T, P, perm_mode=0, seed=12345, nthreads=th)
print("jdk nulls", time.time()-t0, "s", flush=True)
def stats(a, b):
    ok = [t for t in range(T) if not np.isnan(a[t]).all()]
    lr = np.array([np.log(np.nanstd(a[t]) / np.nanstd(b[t])) for t in ok])
    ps = [ks_2samp(a[t][~np.isnan(a[t])], b[t][~np.isnan(b[t])]).pvalue for t in ok]
    # tails: fraction above pooled 99th percentile
    tl = []
    for t in ok:
        q = np.nanpercentile(np.concatenate([a[t], b[t]]), 99)
        tl.append((np.nansum(a[t] > q) + 1) / (np.nansum(b[t] > q) + 1))
    print("   argmin type", ok[int(np.argmin(ps))], "n<0.01:", sum(p < 0.01 for p in ps)); return lr.mean(), lr.std() / np.sqrt(len(ok)), min(ps), float(np.median(ps)), float(np.mean(np.log(tl)))
m, se, lo, med, tl = stats(jdk2, jdk)
print(f"JDK other seed        : mean log sd ratio {m:+.4f} +- {se:.4f}  min p {lo:.3g} med p {med:.3f} tail99 logratio {tl:+.3f}")
for v in variants:
    L.orc_set_mix_variant(C.c_int(v))
    for r in rounds_list:
        for seed in (7, 1234):
            ne, _ = O.null(s, lab, T, P, perm_mode=1, seed=seed, rounds=r, nthreads=th)
            m, se, lo, med, tl = stats(ne, jdk)
            print(f"variant {v} rounds {r} seed {seed:5d}: mean log sd ratio {m:+.4f} +- {se:.4f}  min p {lo:.3g} med p {med:.3f} tail99 logratio {tl:+.3f}", flush=True)
