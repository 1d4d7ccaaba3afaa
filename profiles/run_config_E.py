#!/usr/bin/env python
"""BASELINE configs[4] ("E") as specified: 256 input protein lists against config B's resident 600k-cell matrix, the
lists sharded over the GPUs of one box (contiguous shares, dist.list_slice), -perm 1000 each, every rank calling
pctsea_analyze_batch on its share; the per-list result structs are gathered on every rank (the only exchange, host
plumbing: lists are independent, there is no collective on the data path). One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/run_config_E.py [N_LISTS=256]
    python profiles/run_config_E.py 32            # one GPU, 32 lists

Timing: barrier + synchronize on both sides of the batch calls, MAX over ranks. Rank 0 prints one JSON line and, with
CHECK=k (default 2), compares k lists of its own share and one list of the LAST rank's share with the CPU oracle at
full size (tests/parity.py) — test infrastructure, after the timed region."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import pctsea_parent_b200 as P
from pctsea_parent_b200 import dist as pdist, synth

n_lists = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS["B"]
ds = synth.make_dataset(cfg, dev)
ctx = P.Context(local)
m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes, keepalive=ds)
m.set_labels(ds.label.cpu().numpy(), cfg.n_types)
m.build_gene_index()
first, count = pdist.list_slice(n_lists, rank, world)
lists = [tuple(t.numpy() for t in synth.make_list(cfg, "cpu", s)) for s in range(first, first + count)]
sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
nprm = ctx.null_params(1000, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_FAST, seed=20261019, min_pass=1000)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


# THREADS=k: a pool of k contexts on this GPU, one host thread each, as the reference's web front-end runs its analyses
# on an executor pool (AnalyzeView.java:163): list i goes to context i mod k. The matrix is shared (read-only); every
# context has its own streams and workspaces, so one list's latency-bound ranking and observed-statistics stages run
# beside another list's scoring.
n_threads = max(1, int(os.environ.get("THREADS", "1")))
pool = [ctx] + [P.Context(local) for _ in range(n_threads - 1)]


def run_pool(ls):
    return pdist.analyze_batch_pooled(pool, m, ls, sprm, cfg.min_score, cfg.n_types, nprm)


run_pool(lists[:2 * n_threads])   # warm-up (workspaces, clocks)
barrier()
t0 = time.perf_counter()
out = run_pool(lists)
local_res = torch.from_numpy(out["results"].view(np.uint8).reshape(count, -1).copy()).to(dev)
all_res = pdist.gather_list_results(local_res, n_lists)                    # [n_lists][T * 72 bytes] on every rank
barrier()
wall = time.perf_counter() - t0
tm = ctx.timings()
stats = torch.tensor([wall, float(tm["total_ms"]), float(sum(max(int(n) - 1, 0) for n in out["n_pass"])),
                      float(np.count_nonzero(out["status"]))], dtype=torch.float64, device=dev)
mx, sm = stats.clone(), stats.clone()
if world > 1:
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
line = None
if rank == 0:
    res_all = all_res.cpu().numpy().view(P.TYPE_RESULT_DTYPE).reshape(n_lists, cfg.n_types)
    units = float(sm[2]) * 1000.0
    line = {"workload": f"E: {n_lists} lists x {cfg.list_len} proteins against config B's resident matrix ({cfg.n_cells} cells, nnz {ds.nnz}), "
                        f"{cfg.n_types} types, -perm 1000 each, lists sharded over {world} GPU(s) ({count} on rank 0)",
            "n_gpus": world, "n_lists": n_lists, "contexts_per_gpu": n_threads, "wall_ms": float(mx[0]) * 1e3, "ms_per_list_whole_job": float(mx[0]) * 1e3 / n_lists,
            "device_ms_per_list_on_a_rank": float(mx[1]) / max(count, 1), "cell_permutations_per_s": units / float(mx[0]),
            "lists_failing_an_input_check": int(sm[3]), "significant_types_fdr_0.05_list0": int(np.count_nonzero(res_all[0]["fdr"] <= 0.05)),
            "stage_ms_per_list_rank0": {k: tm[k] / max(count, 1) for k in ("score_ms", "rank_ms", "observed_ms", "null_ms", "reduce_ms")}}
    n_check = int(os.environ.get("CHECK", "2"))
    if n_check > 0:
        from oracle import oracle as O
        from tests import parity as TP
        d = ds.numpy()
        kw = dict(method=cfg.method, min_genes_cells=cfg.min_genes_cells, min_corr=cfg.min_corr, has_min_corr=True, jitter_mode=1, seed=1)
        checked = []
        for gi in list(range(min(n_check, count))) + ([n_lists - 1] if world > 1 else []):
            g, a = (tuple(t.numpy() for t in synth.make_list(cfg, "cpu", gi)))
            try:
                one = ctx.analyze(m, g, a, sprm, cfg.min_score, cfg.n_types, nprm, want_null=True, want_scores=True, want_order=True)
            except P.PctseaError as e:   # a list that fails the reference's 1000-cell guard fails alone in the batch, too
                assert e.code == -4 and np.isnan(res_all[gi]["ews"]).all()
                continue
            assert np.array_equal(one["results"].view(np.uint8), res_all[gi].view(np.uint8)), f"list {gi}: gathered batch result differs from a single analysis"
            TP.check_analysis(O, dict(d, gene_idx=g, abundance=a), cfg.n_genes, cfg.n_types, one, score_params=kw, min_score=cfg.min_score,
                              canonical=2, seed=20261019, perm_indices=(0, 999), sequential_sample=0, what=f"config E list {gi}")
            checked.append(gi)
        line["parity"] = {"ok": True, "lists_checked_vs_oracle_at_full_size": checked}
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if line:
    print(json.dumps(line), flush=True)
