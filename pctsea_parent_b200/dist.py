"""Multi-GPU plumbing: one process per GPU (torch.distributed), permutations sharded by GLOBAL index.

The path has exactly one exchange step: the per-rank null slices [T][P_r] are all-gathered (<= 12 MB in total,
latency-bound over NVLink/NVSwitch) so that every rank can run a10-a12 on the complete null in canonical
permutation order (SURVEY.md §8 e). Works on NCCL (device tensors) and gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def perm_slice(n_perm: int, rank: int, world: int) -> Tuple[int, int]:
    """(perm_begin, perm_count) of `rank`: contiguous, sizes differ by at most one, covers [0, n_perm)."""
    base, rem = divmod(n_perm, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def list_slice(n_lists: int, rank: int, world: int) -> Tuple[int, int]:
    """(first list, number of lists) of `rank` when independent input lists are sharded (BASELINE config E): the
    same contiguous split as perm_slice. Lists need no data-path collective; the per-list results are gathered by
    the host (gather_list_results)."""
    return perm_slice(n_lists, rank, world)


def analyze_batch_pooled(contexts, m, lists, sprm, min_score, n_types, nprm, **kw):
    """Config E on one GPU through a pool of contexts, one host thread each — how the reference's web front-end runs
    its analyses (an executor pool of PCTSEA instances, AnalyzeView.java:163, 747-766): list i goes to context
    i mod len(contexts). The matrix is shared (read-only); every context has its own streams and workspaces, so one
    list's latency-bound ranking and observed-statistics stages run beside another list's scoring (two contexts:
    1.08 -> 0.84 ms per list on config B's matrix; three are slower again). Returns what Context.analyze_batch
    returns, in list order; every list's results are bit-identical to a single-context call."""
    import threading

    import numpy as np
    k = len(contexts)
    if k == 1 or len(lists) < 2:
        return contexts[0].analyze_batch(m, lists, sprm, min_score, n_types, nprm, **kw)
    parts, errors = [None] * k, [None] * k

    def work(j):
        try:
            if lists[j::k]:
                parts[j] = contexts[j].analyze_batch(m, lists[j::k], sprm, min_score, n_types, nprm, **kw)
        except BaseException as e:   # re-raised on the calling thread
            errors[j] = e
    threads = [threading.Thread(target=work, args=(j,)) for j in range(k)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    merged = {}
    first = next(p for p in parts if p is not None)
    for key, val in first.items():
        arr = np.empty((len(lists),) + val.shape[1:], val.dtype)
        for j in range(k):
            if parts[j] is not None:
                arr[j::k] = parts[j][key]
        merged[key] = arr
    return merged


def gather_list_results(local: torch.Tensor, n_lists: int) -> torch.Tensor:
    """local: [lists of this rank][...] -> [n_lists][...] on every rank, in list order (uint8 views of the result
    structs travel as bytes)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == n_lists
        return local.contiguous()
    world = dist.get_world_size()
    counts = [list_slice(n_lists, r, world)[1] for r in range(world)]
    width = max(counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0).contiguous()


def all_gather_null(local: torch.Tensor, n_perm: int) -> torch.Tensor:
    """local: [T][P_rank] fp32 slice of this rank -> [T][n_perm] on every rank, permutations in global order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[1] == n_perm
        return local.contiguous()
    world = dist.get_world_size()
    T = local.shape[0]
    counts = [perm_slice(n_perm, r, world)[1] for r in range(world)]
    width = max(counts)
    if min(counts) == width and local.is_cuda:   # equal slices on NCCL: one collective into [world][T][width], one transpose copy
        out = torch.empty((world, T, width), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out.permute(1, 0, 2).reshape(T, world * width)
    padded = torch.zeros((T, width), dtype=local.dtype, device=local.device)
    padded[:, : local.shape[1]] = local
    parts: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1).contiguous()


def attach_communicator(ctx) -> None:
    """Join `ctx` (this rank's pctsea Context) to an NCCL communicator INSIDE libpctsea_b200.so, using the default
    torch.distributed group only to hand rank 0's unique id to the other ranks (any backend; a Java host would use a
    socket). Afterwards ctx.analyze() is the whole multi-GPU path: permutations sharded over the ranks, null slices
    all-gathered over NVLink on the library's own stream, complete results on every rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init_rank(world, rank, box[0])


class DevicePointer:
    """Zero-copy view of a raw device pointer (the ABI's pctsea_last_null_dev) for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
