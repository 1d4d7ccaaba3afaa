#!/usr/bin/env python
"""How many Feistel rounds does the keyed label permutation need? Null of the enrichment score (oracle, CPU) under the
Feistel permutation with r rounds against the JDK Collections.shuffle null: two-sample KS per cell type.
    python profiles/feistel_rounds_check.py [N_USED=20000] [T=10] [P=20000]
Output of the committed run: profiles/feistel_rounds_check.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import ks_2samp
from oracle import oracle as O
from pctsea_parent_b200 import synth

O.build()
Nu = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
P = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
s, lab = synth.synthetic_ranked(Nu, T, seed=21)
threads = os.cpu_count() or 1
jdk, _ = O.null(s, lab, T, P, perm_mode=0, seed=99, nthreads=threads)
jdk2, _ = O.null(s, lab, T, P, perm_mode=0, seed=12345, nthreads=threads)


def compare(a, b):
    ps = [ks_2samp(a[t][~np.isnan(a[t])], b[t][~np.isnan(b[t])]).pvalue for t in range(T) if not np.isnan(a[t]).all()]
    return min(ps), float(np.median(ps))


print(f"ranking of {Nu} cells, {T} types, {P} permutations per null; smallest / median two-sample KS p over the types")
for r in (2, 3, 4, 6, 8):
    ne, _ = O.null(s, lab, T, P, perm_mode=1, seed=7, rounds=r, nthreads=threads)
    lo, med = compare(ne, jdk)
    print(f"Feistel {r} rounds vs JDK shuffle: min p {lo:.3g}  median p {med:.3f}")
lo, med = compare(jdk2, jdk)
print(f"JDK shuffle (another seed) vs JDK shuffle: min p {lo:.3g}  median p {med:.3f}")
