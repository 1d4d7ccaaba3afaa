#!/usr/bin/env python
"""Regenerates tests/golden/*.json.

The reference (Java + MongoDB) cannot run in this image and ships no numeric fixtures, so the golden vectors
are of three kinds, kept apart in the file:
  published   known answers of third-party algorithms the reference calls (java.util.Random / Collections.shuffle,
              Random123 Philox4x32-10) — independent of this repo;
  survey      the hand-derived bit patterns of SURVEY.md Appendix B.1 (weighted KS, 10 cells, 2 types);
  regression  outputs of the CPU oracle on small seeded inputs, committed so that any later change of the oracle
              or of the CUDA path shows up as a diff (they pin behaviour, not correctness vs Java).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pctsea_parent_b200 import synth  # noqa: E402


def bits32(a):
    return [int(x) for x in np.ascontiguousarray(a, np.float32).view(np.uint32).ravel()]


def bits64(a):
    return [int(x) for x in np.ascontiguousarray(a, np.float64).view(np.uint64).ravel()]


def main():
    out = {}
    out["published"] = {
        "java_random_42_nextInt": [-1170105035, 234785527],
        "java_random_0_nextInt": [-1155484576],
        "java_random_42_nextInt10": [0, 3, 8, 4, 0, 5, 5, 8, 9, 3],
        "philox4x32_10": [
            {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
            {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
            {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
             "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
        ],
    }
    out["survey"] = {
        "b1": {"s": [0.9, 0.85, 0.8, 0.75, 0.7, 0.65, 0.6, 0.55, 0.5, 0.45], "lab": [1, 0, 0, 0, 1, 1, 1, 1, 1, 1],
               "type0": {"raw_sup": 0x3f4b08d4, "supX": 4, "ews": 0x3f1611a8, "dstat": 0x3f931cca, "sizeA": 3, "sizeB": 7},
               "type1": {"raw_sup": 0xbf4b08d4, "supX": 4, "ews": 0xbfcb08d4, "dstat": 0xbf931cca, "sizeA": 7, "sizeB": 3}},
        "shuffle_random42_identity10": [4, 6, 2, 1, 7, 9, 8, 5, 3, 0],
    }
    # regression: a full small pipeline
    cfg = synth.CONFIGS["tiny"]
    d = synth.make_dataset(cfg, "cpu").numpy()
    sc, used, nz, kept = O.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["gene_idx"], d["abundance"],
                                 method=0, min_genes_cells=4, min_corr=0.2, jitter_mode=1, seed=1)
    order, npass = O.rank(sc, kept, 0.0)
    nu = npass - 1
    s_r, lab_r = sc[order[:nu]], d["label"][order[:nu]]
    res = O.ks_observed(s_r, lab_r, cfg.n_types)
    ne_j, nd_j = O.null(s_r, lab_r, cfg.n_types, 8, perm_mode=0, seed=42)
    ne_f, nd_f = O.null(s_r, lab_r, cfg.n_types, 8, perm_mode=1, seed=42)
    red = O.reduce(res, ne_f)
    F = O.fast_frac_bits(s_r)
    fast0 = O.ks_null_fast_one(s_r, O.perm_labels(42, 0, lab_r, cfg.n_types), cfg.n_types, F)
    out["regression"] = {
        "config": "tiny", "n_pass": int(npass), "kept_sum": int(kept.sum()), "n_used_sum": int(used.sum()),
        "score_bits_first32_nonnan": bits64(sc[~np.isnan(sc)][:32]),
        "order_first32": [int(x) for x in order[:32]],
        "ews_bits": bits32(res["ews"]), "dstat_bits": bits32(res["dstat"]), "supX": [int(x) for x in res["supX"]],
        "null_jdk_seed42_bits": bits32(ne_j), "null_feistel_seed42_bits": bits32(ne_f),
        "norm_ews_bits": bits32(red["norm_ews"]), "p_emp_bits": bits64(red["p_emp"]), "fdr_bits": bits64(red["fdr"]),
        "feistel_perm_seed42_p0_first16": [int(x) for x in O.feistel_perm(42, 0, nu)[:16]],
        "fast_frac_bits": int(F), "fast_null_p0_bits": bits32(fast0),
    }
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("written", len(json.dumps(out)), "bytes")


if __name__ == "__main__":
    main()
