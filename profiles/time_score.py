#!/usr/bin/env python
"""Time the scoring stage alone on config B (device-resident CSR): python profiles/time_score.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pctsea_parent_b200 import synth
from pctsea_parent_b200._lib import Context
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = synth.CONFIGS[os.environ.get("CFG", "B")]
ds = synth.make_dataset(cfg, device="cuda")
ctx = Context(0)
m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes, keepalive=ds)
gi, ab = ds.gene_idx.cpu().numpy(), ds.abundance.cpu().numpy()
ts = []
for _ in range(reps + 2):
    ctx.score(m, gi, ab)
    ts.append(ctx.timings()["score_ms"])
print("score_ms", " ".join(f"{t:.3f}" for t in ts[2:]), "min", min(ts[2:]))
