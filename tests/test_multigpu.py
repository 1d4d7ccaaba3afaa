"""The multi-GPU path through the C ABI (pctsea_comm_*): permutations sharded over the ranks, null slices
all-gathered by NCCL inside pctsea_analyze. One-rank communicator on any GPU box; two ranks when two GPUs exist
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

import pctsea_parent_b200 as P
from pctsea_parent_b200 import synth
from tests.helpers import assert_bits_equal_f32, assert_results_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _analysis(ctx, d, cfg, n_perm, ks_mode, **kw):
    m = ctx.load_csr(d["row_ptr"], d["col"], d["val"], cfg.n_genes)
    m.set_labels(d["label"], cfg.n_types)
    m.build_gene_index()
    sprm = P.ScoreParams(0, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
    nprm = ctx.null_params(n_perm, perm_mode=P.PERM_PHILOX, ks_mode=ks_mode, seed=9, min_pass=100, **kw)
    out = ctx.analyze(m, d["gene_idx"], d["abundance"], sprm, cfg.min_score, cfg.n_types, nprm, want_null=True,
                      want_scores=True, want_order=True)
    tm = ctx.timings()
    m.free()
    return out, tm


@pytest.mark.parametrize("ks_mode", [P.KS_FAST, P.KS_EXACT])
def test_one_rank_communicator_is_invisible(ctx, small, ks_mode):
    """nranks = 1: NCCL is loaded, the all-gather and the reorder kernel run, and every output equals the call
    without a communicator bit for bit."""
    cfg = small["cfg"]
    ref, _ = _analysis(ctx, small, cfg, 37, ks_mode)
    c2 = P.Context(0)
    try:
        c2.comm_init_rank(1, 0, P.Context.comm_unique_id())
        out, tm = _analysis(c2, small, cfg, 37, ks_mode)
        assert tm["gather_ms"] > 0.0
        assert_bits_equal_f32(out["null"], ref["null"], "null")
        assert_bits_equal_f32(out["nullD"], ref["nullD"], "null dStat")
        assert_results_equal(out["results"], ref["results"], None, "one-rank communicator")
        with pytest.raises(P.PctseaError):   # a communicator call covers all permutations
            _analysis(c2, small, cfg, 37, ks_mode, perm_begin=3, perm_count=5)
        c2.comm_destroy()
        out2, tm2 = _analysis(c2, small, cfg, 37, ks_mode)
        assert tm2["gather_ms"] == 0.0
        assert_bits_equal_f32(out2["null"], ref["null"], "null after comm_destroy")
    finally:
        c2.close()


def _rank_main(rank, world, uid_q, res_q, n_perm):
    sys.path.insert(0, ROOT)
    import pctsea_parent_b200 as PP
    from pctsea_parent_b200 import synth as S
    cfg = S.CONFIGS["small"]
    d = S.make_dataset(cfg, "cpu").numpy()
    ctx = PP.Context(rank)
    if rank == 0:
        uid = PP.Context.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    ctx.comm_init_rank(world, rank, uid)
    out, tm = _analysis(ctx, d, cfg, n_perm, PP.KS_FAST)
    # the observed statistics are sharded by cell type: the a / b series of a type this rank did NOT walk (the other
    # rank's first type) and of one it did
    n_used = int(tm["n_used"])
    other = (cfg.n_types + world - 1) // world * ((rank + 1) % world)
    curves = {t: ctx.last_observed_curve(t, n_used) for t in (other, min(other + 1, cfg.n_types - 1), 0)}
    res_q.put((rank, out["null"], out["nullD"], out["results"], tm["gather_ms"], out["score"], out["n_genes_used"], out["kept"],
               out["order"], curves))
    ctx.close()


@pytest.mark.parametrize("n_perm", [64, 37])   # equal and unequal slices
def test_two_rank_analysis_equals_single_gpu(ctx, small, n_perm):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = small["cfg"]
    ref, rtm = _analysis(ctx, small, cfg, n_perm, P.KS_FAST)
    ref_curves = {t: ctx.last_observed_curve(t, int(rtm["n_used"])) for t in range(cfg.n_types)}
    sp = mp.get_context("spawn")
    uid_q, res_q = sp.Queue(), sp.Queue()
    procs = [sp.Process(target=_rank_main, args=(r, 2, uid_q, res_q, n_perm)) for r in range(2)]
    for p in procs:
        p.start()
    got = [res_q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, null, nullD, res, gather_ms, score, used, kept, order, curves in got:
        for t, (a, b) in curves.items():
            assert_bits_equal_f32(a, ref_curves[t][0], f"rank {rank}: observed curve a of type {t}")
            assert_bits_equal_f32(b, ref_curves[t][1], f"rank {rank}: observed curve b of type {t}")
        assert gather_ms > 0.0
        # stage 1 was sharded over the two ranks (every second tile of 256 cells each) and all-gathered
        assert np.array_equal(score.view(np.uint64), ref["score"].view(np.uint64)), f"rank {rank}: scores"
        assert np.array_equal(used, ref["n_genes_used"]) and np.array_equal(kept, ref["kept"]) and np.array_equal(order, ref["order"])
        assert_bits_equal_f32(null, ref["null"], f"rank {rank}: pooled null")
        assert_bits_equal_f32(nullD, ref["nullD"], f"rank {rank}: pooled null dStat")
        assert_results_equal(res, ref["results"], None, f"rank {rank}")
