"""world_size-2 gloo test of the multi-GPU plumbing on CPU: permutations sharded by global index, null slices
all-gathered in permutation order, a10-a12 on the gathered null equal the single-process result. The per-rank
compute is the oracle here (tests may use it); on GPUs the same plumbing moves device tensors over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pctsea_parent_b200 import dist as pdist
from pctsea_parent_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_perm, out_dir):
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, lab = synth.synthetic_ranked(1500, 7, seed=3)
    T = 7
    begin, count = pdist.perm_slice(n_perm, rank, world)
    ne, _ = O.null(s, lab, T, count, perm_mode=1, seed=77, perm_begin=begin)
    full = pdist.all_gather_null(torch.from_numpy(ne), n_perm)
    res = O.reduce(O.ks_observed(s, lab, T), full.numpy())
    np.save(os.path.join(out_dir, f"null_{rank}.npy"), full.numpy())
    np.save(os.path.join(out_dir, f"fdr_{rank}.npy"), res["fdr"])
    dist.barrier()
    dist.destroy_process_group()


def test_perm_slice_covers_everything():
    for n in (0, 1, 7, 1000, 10_001):
        for w in (1, 2, 3, 8):
            sl = [pdist.perm_slice(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sum(c for _, c in sl) == n
            for (b0, c0), (b1, _) in zip(sl, sl[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in sl) - min(c for _, c in sl) <= 1


@pytest.mark.parametrize("n_perm", [10, 11])
def test_two_rank_null_gather_matches_single_process(tmp_path, n_perm):
    from oracle import oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_perm, str(tmp_path)), nprocs=world, join=True)
    s, lab = synth.synthetic_ranked(1500, 7, seed=3)
    ref, _ = O.null(s, lab, 7, n_perm, perm_mode=1, seed=77)
    ref_fdr = O.reduce(O.ks_observed(s, lab, 7), ref)["fdr"]
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"null_{r}.npy"))
        assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        fdr = np.load(os.path.join(tmp_path, f"fdr_{r}.npy"))
        assert np.array_equal(np.isnan(fdr), np.isnan(ref_fdr)) and np.array_equal(fdr[~np.isnan(fdr)], ref_fdr[~np.isnan(ref_fdr)])


def _list_worker(rank, world, port, n_lists, out_dir):
    """Config E plumbing: independent lists sharded over ranks, per-list result structs gathered as bytes."""
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 5
    first, count = pdist.list_slice(n_lists, rank, world)
    rows = []
    for i in range(first, first + count):
        s, lab = synth.synthetic_ranked(1200, T, seed=100 + i)
        ne, _ = O.null(s, lab, T, 8, perm_mode=1, seed=5)
        rows.append(O.reduce(O.ks_observed(s, lab, T), ne))
    local = np.stack(rows) if rows else np.zeros((0, T), O.TYPE_RESULT_DTYPE)
    as_bytes = torch.from_numpy(local.view(np.uint8).reshape(local.shape[0], -1).copy())
    full = pdist.gather_list_results(as_bytes, n_lists)
    np.save(os.path.join(out_dir, f"lists_{rank}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_lists", [4, 5])
def test_two_rank_list_sharding_matches_single_process(tmp_path, n_lists):
    from oracle import oracle as O
    world, T = 2, 5
    mp.spawn(_list_worker, args=(world, _free_port(), n_lists, str(tmp_path)), nprocs=world, join=True)
    rows = []
    for i in range(n_lists):
        s, lab = synth.synthetic_ranked(1200, T, seed=100 + i)
        ne, _ = O.null(s, lab, T, 8, perm_mode=1, seed=5)
        rows.append(O.reduce(O.ks_observed(s, lab, T), ne))
    ref = np.stack(rows)
    ref_bytes = ref.view(np.uint8).reshape(n_lists, -1)
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"lists_{r}.npy"))
        assert got.shape == ref_bytes.shape and np.array_equal(got, ref_bytes)


def test_list_slice_covers_everything():
    for n in (0, 1, 31, 256):
        for w in (1, 2, 8):
            sl = [pdist.list_slice(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sum(c for _, c in sl) == n


def test_pooled_batch_merges_in_list_order_and_propagates_errors():
    """dist.analyze_batch_pooled (config E through a pool of contexts, one host thread each): list i goes to context
    i mod k, the per-list outputs come back in list order, an exception of one context reaches the caller. Host logic
    only: the contexts are stand-ins (the GPU test runs real ones)."""
    import numpy as np
    import pytest
    from pctsea_parent_b200 import dist as pdist

    class Fake:
        def __init__(self, tag, fail=False):
            self.tag, self.fail, self.seen = tag, fail, []

        def analyze_batch(self, m, lists, sprm, min_score, n_types, nprm, **kw):
            if self.fail:
                raise RuntimeError("boom")
            self.seen.append([int(g[0]) for g, _ in lists])
            ids = np.array([int(g[0]) for g, _ in lists])
            return {"results": np.stack([np.full(n_types, 10 * i + self.tag) for i in ids]), "n_pass": ids + 1000,
                    "status": np.zeros(len(lists), np.int32)}

    lists = [(np.array([i], np.int32), np.ones(1, np.float32)) for i in range(7)]
    a, b = Fake(1), Fake(2)
    out = pdist.analyze_batch_pooled([a, b], None, lists, None, 0.0, 3, None)
    assert a.seen == [[0, 2, 4, 6]] and b.seen == [[1, 3, 5]]
    assert list(out["n_pass"]) == [1000 + i for i in range(7)]
    assert [int(r[0]) for r in out["results"]] == [10 * i + (1 if i % 2 == 0 else 2) for i in range(7)]
    one = pdist.analyze_batch_pooled([Fake(5)], None, lists, None, 0.0, 3, None)
    assert list(one["n_pass"]) == [1000 + i for i in range(7)]
    with pytest.raises(RuntimeError):
        pdist.analyze_batch_pooled([Fake(1), Fake(2, fail=True)], None, lists, None, 0.0, 3, None)
