#!/usr/bin/env python
"""BASELINE config E on one GPU's share: B's matrix resident, N_LISTS lists (list seeds FIRST..FIRST+N_LISTS-1) of
1000 proteins, -perm 1000 each, through pctsea_analyze_batch; beside it the same lists one pctsea_analyze call at
a time. Prints one JSON line.   python profiles/time_batch.py [N_LISTS=32] [FIRST=0] [CONFIG=B]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth

n_lists = int(sys.argv[1]) if len(sys.argv) > 1 else 32
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = synth.CONFIGS[sys.argv[3] if len(sys.argv) > 3 else "B"]
dev = torch.device("cuda", 0)
ds = synth.make_dataset(cfg, dev)
ctx = P.Context(0)
m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes, keepalive=ds)
m.set_labels(ds.label.cpu().numpy(), cfg.n_types)
if os.environ.get("PCTSEA_BENCH_GENE_INDEX", "1") != "0":
    m.build_gene_index()
lists = []
for s in range(first, first + n_lists):
    g, a = synth.make_list(cfg, "cpu", s)
    lists.append((g.numpy(), a.numpy()))
sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
nprm = ctx.null_params(1000, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=20261019, min_pass=1000)

ctx.analyze_batch(m, lists[:2], sprm, cfg.min_score, cfg.n_types, nprm)   # warm-up (workspaces, clocks)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = ctx.analyze_batch(m, lists, sprm, cfg.min_score, cfg.n_types, nprm)
torch.cuda.synchronize()
wall_batch = time.perf_counter() - t0
tm = ctx.timings()
t0 = time.perf_counter()
singles = [ctx.analyze(m, g, a, sprm, cfg.min_score, cfg.n_types, nprm) for g, a in lists]
torch.cuda.synchronize()
wall_single = time.perf_counter() - t0
same = all(np.array_equal(out["results"][i].view(np.uint8), singles[i]["results"].view(np.uint8)) for i in range(n_lists))
units = float(sum(max(int(n) - 1, 0) for n in out["n_pass"])) * 1000.0
print(json.dumps({
    "score_path": "gene-major index" if m.gene_index else "csr-scan",
    "workload": f"E share: {n_lists} lists x {cfg.list_len} proteins against config {cfg.name}'s resident matrix "
                f"({cfg.n_cells} cells, nnz {ds.nnz}), {cfg.n_types} types, -perm 1000 each, 1 GPU",
    "batch_wall_ms": wall_batch * 1e3, "batch_ms_per_list": wall_batch * 1e3 / n_lists,
    "batch_device_ms_per_list": tm["total_ms"] / n_lists,
    "single_calls_wall_ms_per_list": wall_single * 1e3 / n_lists,
    "cell_permutations_per_s": units / wall_batch,
    "stage_ms_per_list": {k: tm[k] / n_lists for k in ("score_ms", "rank_ms", "observed_ms", "labels_ms", "null_ms", "reduce_ms")},
    "n_pass_min_max": [int(out["n_pass"].min()), int(out["n_pass"].max())],
    "status_nonzero": int(np.count_nonzero(out["status"])), "launches": int(tm["launches"]),
    "batch_bit_identical_to_single_calls": bool(same)}))
