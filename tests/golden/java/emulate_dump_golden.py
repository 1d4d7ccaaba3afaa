"""What tests/golden/java/DumpGolden.java writes, computed with the ORACLE in place of the reference classes — same
inputs (java.util.Random restated), same JSON layout. Its only use is to test the consumer
(tests/test_oracle_cpu.py::test_golden_ref_consumer_on_emulated_dump) in an image without a JDK. Its output pins
nothing and must never be committed as tests/golden/golden_ref.json."""
import json
import math
import struct
import sys

import numpy as np


def f64(v):
    return format(struct.unpack("<Q", struct.pack("<d", float(v)))[0], "x")


def f32(v):
    return format(struct.unpack("<I", struct.pack("<f", np.float32(v)))[0], "x")


def java_round(x):   # Math.round(double) = floor(x + 0.5)
    return math.floor(x + 0.5)


def emulate(O, path):
    G = {"generator": "tests/golden/java/emulate_dump_golden.py (oracle, NOT the reference)", "java": None}
    bounds = [10, 16, 1000, 219530, 1 << 30, (1 << 30) + 1]
    G["jdk_random"] = []
    for seed in (42, 0):
        r = O.JRandom(seed)
        G["jdk_random"].append({"seed": seed, "nextInt": [r.next_int() for _ in range(8)], "nextIntBounds": bounds,
                                "nextIntBounded": [r.next_int(b) for b in bounds],
                                "nextDouble": [f64(r.next_double()) for _ in range(4)]})
    G["shuffle"] = [{"n": n, "seed": 42, "result": O.JRandom(42).shuffle(np.arange(n, dtype=np.int32)).tolist()}
                    for n in (10, 257)]
    r = O.JRandom(3)
    G["pearson"] = []
    for n in (2, 3, 5, 17, 120):
        x, y = [], []
        for _ in range(n):
            x.append(float(1 + r.next_int(40)))
            y.append(float(1 + r.next_int(12)))
        G["pearson"].append({"x": [f64(v) for v in x], "y": [f64(v) for v in y], "r": f64(O.pearson(x, y))})
    a = np.array([0.5, -0.0, 0.0, 2.0, 0.5, 0.5, np.nan, 1.25, 0.5, -3.0, 2.0, 0.5], np.float32)
    s = O.java_sort_f32(a)
    keys = np.array([0.5, 2.0, 0.0, -0.0, 1.0, 9.0, -9.0, np.nan], np.float32)
    G["sort_search"] = {"input": [f32(v) for v in a], "sorted": [f32(v) for v in s],
                        "searches": [{"key": f32(k), "index": int(O.binary_search_f32(s, float(k)))} for k in keys]}
    n, T = 400, 5
    rg = O.JRandom(11)
    score, typ = np.empty(n), np.empty(n, np.int32)
    for i in range(n):
        score[i] = 0.2 + 0.8 * java_round(rg.next_double() * 64) / 64.0
        t = rg.next_int(T + 1)
        typ[i] = -1 if t == T else (0 if rg.next_int(4) == 0 else t)
    kept = np.ones(n, np.uint8)
    order, npass = O.rank(score, kept, 0.0)   # = sortByScoreDescending here: every score passes; score desc, stable (index asc)
    assert npass == n
    s_r, lab_r = score[order], typ[order]
    G["ks_statistic"] = {"type": 0, "d": f64(O.ks_statistic(s_r[lab_r == 0], s_r[lab_r != 0]))}
    obs = O.ks_observed(s_r, lab_r, T)
    observed = []
    for t in range(T):
        has_sec = int(obs["sec_supX"][t]) != 0
        observed.append({"type": t, "ews": f32(obs["ews"][t]), "supX": int(obs["supX"][t]), "norm_supX": f64(obs["norm_supX"][t]),
                         "sizeA": int(obs["sizeA"][t]), "sizeB": int(obs["sizeB"][t]), "dstat": f32(obs["dstat"][t]),
                         "sec_ews": f32(obs["sec_ews"][t]) if has_sec else None,
                         "sec_supX": int(obs["sec_supX"][t]) if has_sec else None, "ks_d": f64(obs["ks_d"][t])})
    rp = O.JRandom(7)
    P = 6
    perm_labels, null, nulld = [], np.empty((T, P), np.float32), np.empty((T, P), np.float32)
    for p in range(P):
        idx = rp.shuffle(np.arange(n, dtype=np.int32))   # shuffling a copy of the label list = indexing by the shuffled identity
        labp = lab_r[idx]
        perm_labels.append([str(int(v)) for v in labp])
        null[:, p], nulld[:, p] = O.ks_null_one(s_r, labp, T)
    red = O.reduce(obs, null)
    G["enrichment"] = {"n_types": T, "score": [f64(v) for v in s_r], "type": [str(int(v)) for v in lab_r],
                       "observed": observed, "perm_seed": 7, "perm_labels": perm_labels,
                       "null": [{"ews": [f32(v) for v in null[t]], "dstat": [f32(v) for v in nulld[t]]} for t in range(T)],
                       "normalized": [f32(v) for v in red["norm_ews"]]}
    json.dump(G, open(path, "w"))


if __name__ == "__main__":
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
    from oracle import oracle as O
    O.build()
    emulate(O, sys.argv[1])
