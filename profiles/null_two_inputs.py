#!/usr/bin/env python
"""The fused null kernel on two inputs of the same size: config B's real ranked list (scores and labels as the analysis
produces them) and the synthetic ranked list profiles/time_null.py uses. Observed kernels kept off the GPU
(OBS_OVERLAP 0). Run plain for the CUDA-event times, or under
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum -k regex:k_null_fused
to compare instruction and shared-memory wavefront counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth

cfg = synth.CONFIGS["B"]
dev = torch.device("cuda", 0)
ds = synth.make_dataset(cfg, dev)
ctx = P.Context(0)
ctx.set_option(P.OPT_OBS_OVERLAP, 0)
m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes, keepalive=ds)
label = ds.label.cpu().numpy()
m.set_labels(label, cfg.n_types)
m.build_gene_index()
gi, ab = ds.gene_idx.cpu().numpy(), ds.abundance.cpu().numpy()
sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
nprm = ctx.null_params(1000, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=20261019, min_pass=1000)
for _ in range(3):
    out = ctx.analyze(m, gi, ab, sprm, cfg.min_score, cfg.n_types, nprm, want_scores=True, want_order=True)
    t = ctx.timings()
print(f"analyze (real data):     null {t['null_ms']:.3f} ms  observed {t['observed_ms']:.3f} ms", flush=True)
nu = out["n_pass"] - 1
order = out["order"]
for _ in range(3):
    ctx.enrich(out["score"], order[:nu], label, cfg.n_types, nprm, want_null=False)
    t = ctx.timings()
print(f"enrich  (real ranking):  null {t['null_ms']:.3f} ms  observed {t['observed_ms']:.3f} ms", flush=True)
lab_r = label[order[:nu]]
cnt = np.bincount(lab_r[lab_r >= 0], minlength=cfg.n_types)
print("real ranking: class sizes min/median/max", cnt.min(), int(np.median(cnt)), cnt.max(), "unknown", int((lab_r < 0).sum()),
      "score range", float(out["score"][order[nu - 1]]), float(out["score"][order[0]]), flush=True)
s, lab = synth.synthetic_ranked(nu + 1, cfg.n_types, seed=1)
o2 = np.arange(nu + 1, dtype=np.int32)
for _ in range(3):
    ctx.enrich(s, o2[:nu], lab, cfg.n_types, nprm, want_null=False)
    t = ctx.timings()
cnt = np.bincount(lab[:nu][lab[:nu] >= 0], minlength=cfg.n_types)
print(f"enrich  (synthetic):     null {t['null_ms']:.3f} ms  observed {t['observed_ms']:.3f} ms; class sizes min/median/max",
      cnt.min(), int(np.median(cnt)), cnt.max(), flush=True)
# synthetic scores with the real labels, and the other way round
for name, sc, lb in (("synthetic scores + real labels", s[:nu], lab_r), ("real scores + synthetic labels", out["score"][order[:nu]], lab[:nu])):
    for _ in range(3):
        ctx.enrich(np.ascontiguousarray(sc), o2[:nu], np.ascontiguousarray(lb), cfg.n_types, nprm, want_null=False)
        t = ctx.timings()
    print(f"enrich  ({name}): null {t['null_ms']:.3f} ms", flush=True)
