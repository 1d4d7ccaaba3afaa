#!/usr/bin/env python
"""bench.py — cells x permutations KS-scored per second on the BASELINE.json workload.

One step = one full PCTSEA analysis (score -> rank -> observed KS -> permutation null -> p/norm/FDR) of one input
list against the resident matrix. N=1 runs BASELINE configs[1] ("B": 600k cells x 25k genes ~5 %, 100 cell
types, 1000-protein list, -perm 1000). Under torchrun each rank owns one GPU with a replica of the matrix and
computes a slice of the permutations (weak scaling: 1000 permutations per GPU, i.e. B -> C as N grows); the
null slices are all-gathered over NCCL and reduced.

  value  : Nu * P_total / (device time of the step's stages, inputs resident in HBM), max over ranks
  e2e    : the same through the C ABI with HOST buffers (list in, per-type results out), wall clock around the
           K steps bracketed by barrier + synchronize, max over ranks
  --impl reference : the CPU restatement of the reference's algorithm (oracle/), all host threads, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UNIT = 10.0  # 8 B fp64 score + 2 B label per (ranked cell x permutation)


def profiled_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu --set full
    capture of this same command (profiles/r2_traffic.json); None if that kernel was not captured."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["kernels"]
        return float(d[kernel]["dram_bytes"]) if kernel in d else None
    except Exception:
        return None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.interval_s = float(os.environ.get("PCTSEA_BENCH_SAMPLE_MS", "25")) * 1e-3
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        self._names = names
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(self.interval_s)

    def _sample(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for k, bit in getattr(self, "_names", {}).items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def start(self):
        if self.nv is not None:
            # the first NVML queries can take ~100 ms and hold driver locks: pay for them here, not inside a step
            self._sample()
            self.samples.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def workload_string(cfg, P_per_gpu, P_total):
    return (f"{cfg.name}: {cfg.n_cells} cells x {cfg.n_genes} genes (~{cfg.density:.0%}), {cfg.n_types} cell types, "
            f"{cfg.list_len}-protein list, {'PEARSONS_CORRELATION' if cfg.method == 0 else 'SIMPLE_SCORE'}, "
            f"-perm {P_per_gpu} per GPU ({P_total} total), min_genes_cells {cfg.min_genes_cells}, min_corr {cfg.min_corr}")


def workload_config(name: str, world: int):
    from pctsea_parent_b200 import synth
    cfg = synth.CONFIGS[name]
    return cfg


def run_ours(args):
    import torch
    import torch.distributed as dist

    import pctsea_parent_b200 as P
    from pctsea_parent_b200 import dist as pdist
    from pctsea_parent_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pctsea_parent_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload_config(args.config, world)
    P_per_gpu = args.perm if args.perm else cfg.n_perm
    P_total = P_per_gpu * world
    ks_mode = P.KS_EXACT if args.ks == "exact" else P.KS_FAST

    t0 = time.time()
    ds = synth.make_dataset(cfg, dev)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    ctx = P.Context(local)
    m = ctx.wrap_device_csr(ds.row_ptr.data_ptr(), ds.col.data_ptr(), ds.val.data_ptr(), cfg.n_cells, cfg.n_genes,
                            keepalive=ds)
    label_host = ds.label.cpu().numpy()
    m.set_labels(label_host, cfg.n_types)
    index_s = None
    if args.gene_index:   # once per dataset, like the matrix load itself: outside the timed region
        t0 = time.time()
        m.build_gene_index()
        index_s = time.time() - t0
    gene_idx = ds.gene_idx.cpu().numpy()
    abundance = ds.abundance.cpu().numpy()
    sprm = P.ScoreParams(cfg.method, cfg.min_genes_cells, cfg.min_corr, 1, P.JITTER_PHILOX, 1)
    if world > 1:
        # NCCL communicator inside the library (pctsea_comm_init_rank): one analyze call is then the whole multi-GPU
        # path — permutations sharded over the ranks, null slices all-gathered on the library's stream, a10-a12 on
        # the complete null; torch.distributed only carries rank 0's unique id to the other ranks
        pdist.attach_communicator(ctx)
    nprm = ctx.null_params(P_total, perm_mode=P.PERM_PHILOX, ks_mode=ks_mode, seed=20261019, min_pass=1000,
                           feistel_rounds=args.feistel_rounds)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(prm=None):
        out = ctx.analyze(m, gene_idx, abundance, sprm, cfg.min_score, cfg.n_types, prm or nprm)
        return out["results"], ctx.timings()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler (NVML thread) starts before the warm-up so that its start-up cost is not inside the timed
    # region; it keeps sampling through the timed steps
    sampler = ClockSampler(local)
    if os.environ.get("PCTSEA_BENCH_NO_SAMPLER"):   # diagnosis only: the JSON line then carries no clocks
        sampler.nv = None
    sampler.start()
    for _ in range(max(args.warmup, 0)):
        flush.zero_()
        step()
    barrier()
    stage_ms = []
    tms = []
    t_begin = time.perf_counter()
    flush_s = 0.0
    for _ in range(args.steps):
        tf = time.perf_counter()
        flush.zero_()          # L2 flush between timed iterations (excluded from the step time below)
        torch.cuda.synchronize()
        flush_s += time.perf_counter() - tf
        res, tm = step()
        tms.append(tm)
        # device time of the whole analysis (begin -> end events on the library's stream; the observed-statistics
        # kernels overlap the null stage on a second stream, so the stage times do not add up to it)
        stage_ms.append(tm["total_ms"] - tm["setup_ms"])
    barrier()
    wall_s = time.perf_counter() - t_begin - flush_s
    clocks = sampler.stop()

    # ---- strong scaling (BASELINE configs[2], "C"): ONE analysis with -perm 10000 on the same matrix, its permutations
    # sharded over the N GPUs, timed end to end in the same run
    strong = None
    if args.strong_perm > 0 and cfg.name == "B":
        prm_s = ctx.null_params(args.strong_perm, perm_mode=P.PERM_PHILOX, ks_mode=ks_mode, seed=20261019, min_pass=1000,
                                feistel_rounds=args.feistel_rounds)
        ks = max(3, min(args.steps, 10))

        def timed_strong():
            for _ in range(2):
                step(prm_s)
            barrier()
            t0 = time.perf_counter()
            dev_t = []
            for _ in range(ks):
                _, tm_s = step(prm_s)
                dev_t.append(tm_s["total_ms"] - tm_s["setup_ms"])
            barrier()
            return (time.perf_counter() - t0) * 1e3 / ks, float(np.mean(dev_t)), tm_s

        s_e2e, s_dev, tm_s = timed_strong()
        if world > 1:
            t = torch.tensor([s_e2e, s_dev], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_e2e, s_dev = float(t[0]), float(t[1])
        strong = {"config": "C", "P_total": args.strong_perm, "n_gpus": world, "steps": ks, "ms_e2e": s_e2e,
                  "ms_device": s_dev,
                  "stage_ms": {k: float(tm_s[k]) for k in ("score_ms", "rank_ms", "observed_ms", "null_ms", "gather_ms", "reduce_ms")}}

    # ---- the dominant kernel alone: in the timed steps the observed-statistics kernels run beside the null kernel on a
    # second stream (that is what makes the step short), so the CUDA events around the null launch also span their
    # interference. For the roofline the same steps are repeated with the observed kernels kept on the main stream
    # (PCTSEA_OPT_OBS_OVERLAP = 0): the events then bracket k_null_fused and nothing else.
    iso_null_ms, iso_total_ms = [], []
    ctx.set_option(P.OPT_OBS_OVERLAP, 0)
    try:
        for i in range(3 + min(args.steps, 10)):
            flush.zero_()
            _, tm_i = step()
            if i >= 3:
                iso_null_ms.append(tm_i["null_ms"])
                iso_total_ms.append(tm_i["total_ms"] - tm_i["setup_ms"])
    finally:
        ctx.set_option(P.OPT_OBS_OVERLAP, 2)
    barrier()

    parity = None if args.no_parity else parity_gate(ctx, m, ds, cfg, sprm, nprm, gene_idx, abundance, args, rank, world,
                                                     P_per_gpu, P_total)

    nu = tms[-1]["n_used"]
    dev_ms = float(np.mean(stage_ms))
    e2e_ms = wall_s * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
    units = float(nu) * float(P_total)

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        mean = lambda k: float(np.mean([t[k] for t in tms]))
        lab_ms, null_ms = mean("labels_ms"), mean("null_ms")
        ks_name = "k_null_fused" if ks_mode == P.KS_FAST else "k_ks_exact"
        # the kernel with the largest share of a step (profiles/r2_launches.txt): the fused null kernel, which generates
        # the permuted labels and evaluates them (the label kernel and the KS kernel of round 1 in one)
        kernel_ms = float(np.mean(iso_null_ms)) + (0.0 if ks_mode == P.KS_FAST else lab_ms)
        units_gpu = float(nu) * float(P_per_gpu)
        ach = ALGO_BYTES_PER_UNIT * units_gpu / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        nnz = ds.nnz
        score_ms = mean("score_ms")
        score_dram = profiled_traffic("k_score_gm" if args.gene_index else "k_score") if cfg.name == "B" else None
        line = {
            "metric": "cells x permutations KS-scored per second", "value": units / (dev_ms * 1e-3),
            "unit": "cell-permutations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": e2e_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 scores / int32 fixed-point running sums / f32 enrichment" if ks_mode == P.KS_FAST else "f64 scores / f32 running sums",
            "data": "synthetic",
            "config": {"workload": workload_string(cfg, P_per_gpu, P_total),
                       "nnz": nnz, "Np": int(tms[-1]["n_pass"]), "Nu": int(nu), "n_perm_total": P_total,
                       "perm_mode": "PHILOX", "ks_mode": args.ks.upper(), "parallelism": f"perm-shard x{world}" + (" + stage-1 cell-shard" if world > 1 and args.gene_index else ""),
                       "score_path": "gene-major index" if args.gene_index else "csr-scan",
                       "l2": "256 MiB flush between timed steps; CSR inputs (>= 6 GB) exceed the 126 MB L2"},
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "cell-permutations/s",
                    "h2d_bytes_per_step": int(mean("h2d_bytes")), "d2h_bytes_per_step": int(mean("d2h_bytes")),
                    "analysis_seconds": e2e_ms * 1e-3},
            "gpu_launches": int(sum(t["launches"] for t in tms)),
            "stage_ms": {k: mean(k) for k in ("setup_ms", "score_ms", "rank_ms", "observed_ms", "labels_ms", "null_ms", "gather_ms", "reduce_ms", "total_ms")},
            "roofline": {"bound": "hbm", "kernel": ks_name, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak,
                         "traffic": profiled_traffic(ks_name) if cfg.name == "B" and P_per_gpu == 1000 else None,
                         "traffic_note": "DRAM bytes of one launch (ncu --set full, config B, profiles/r2_traffic.json): the kernel's "
                                         "inputs (fixed-point scores in lane layout, class tables) and its per-CTA label scratch "
                                         "are L2-resident by design, so DRAM traffic is far below the notional 10 B/unit stream; "
                                         "the kernel is bound by instruction issue (ALU pipe), not by HBM",
                         "algorithmic_bytes": ALGO_BYTES_PER_UNIT * units_gpu, "peak_source": peak_src,
                         "kernel_ms": kernel_ms,
                         "kernel_ms_note": "CUDA events around the kernel's launch on the library stream, observed-statistics "
                                           "kernels kept off the GPU meanwhile (separate pass after the timed steps, same inputs)",
                         "kernel_ms_in_step": null_ms + (0.0 if ks_mode == P.KS_FAST else lab_ms),
                         "step_ms_without_overlap": float(np.mean(iso_total_ms)),
                         "null_stage_frac": ach / peak,
                         "score": {"kernel": "k_score_gm" if args.gene_index else "k_score", "stage_ms": score_ms,
                                   "dram_bytes": score_dram,
                                   "dram_gbs": score_dram / (score_ms * 1e-3) / 1e9 if score_dram and score_ms > 0 else None,
                                   "dram_frac_of_peak": score_dram / (score_ms * 1e-3) / 1e9 / peak if score_dram and score_ms > 0 else None,
                                   "csr_bytes_not_read": 8.0 * nnz,
                                   "note": "the gene-major kernel reads the list's genes only (entries + short value runs: every "
                                           "run costs at least a 32-byte sector), not the CSR; dram_bytes is the ncu measurement of "
                                           "one launch (stage time here also covers the list lookup and the final-score kernels)"}},
            "clocks": clocks,
            "per_step_total_ms": [round(t["total_ms"], 3) for t in tms],
            "setup": {"generate_s": gen_s, "gene_index_s": index_s},
        }
        line["parity"] = parity
        line["strong"] = strong
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(ctx, m, ds, cfg, sprm, gene_idx, abundance, label_host, args)
    if strong is not None:
        # the 1-GPU time of the same job, for the speed-up: rank 0 alone, without the communicator
        if world > 1:
            ctx.comm_destroy()
            dist.barrier()
        if rank == 0:
            if world > 1:
                one_e2e, one_dev, _ = timed_strong_single(ctx, step, prm_s, ks)
            else:
                one_e2e, one_dev = strong["ms_e2e"], strong["ms_device"]
            strong["ms_e2e_1gpu"], strong["ms_device_1gpu"] = one_e2e, one_dev
            strong["speedup_vs_1gpu"] = one_e2e / strong["ms_e2e"]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    m.free()
    ctx.close()
    if parity is not None and not parity.get("ok", False):
        sys.stderr.write(f"bench.py: PARITY GATE FAILED: {parity.get('failed')}\n")
        raise SystemExit(3)


def timed_strong_single(ctx, step, prm_s, ks):
    import torch
    for _ in range(2):
        step(prm_s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_t = []
    for _ in range(ks):
        _, tm_s = step(prm_s)
        dev_t.append(tm_s["total_ms"] - tm_s["setup_ms"])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / ks, float(np.mean(dev_t)), tm_s


def parity_gate(ctx, m, ds, cfg, sprm, nprm, gene_idx, abundance, args, rank, world, P_per_gpu, P_total):
    """no number without parity. After the timed region the same analysis is run once more with its
    scores, ranking and null copied out, and rank 0 compares them with the CPU oracle at FULL size (tests/parity.py:
    all scores bit for bit, the ranking, every observed field, three permutations of its own null slice bit for bit,
    a10-a12 from the GPU's null). Under --gpus N the null is the pooled one the library all-gathered, and the gate also
    recomputes one permutation of EVERY other rank's slice of it. The oracle is the checker here,
    never the thing measured. Returns the report that goes into the JSON line; a mismatch makes bench.py exit 3."""
    out = ctx.analyze(m, gene_idx, abundance, sprm, cfg.min_score, cfg.n_types, nprm, want_null=True,
                      want_scores=True, want_order=True)   # under --gpus N: the pooled null, on every rank
    if rank != 0:
        return None
    from oracle import oracle as O
    from tests import parity as TP
    d = ds.numpy()
    fast = args.ks != "exact"
    # three permutations of rank 0's share and one of every other rank's share of the global permutation indices
    perms = sorted({0, P_per_gpu // 2, P_per_gpu - 1} | {r * P_per_gpu + (131 * r) % P_per_gpu for r in range(1, world)})
    score_kw = dict(method=cfg.method, min_genes_cells=cfg.min_genes_cells, min_corr=cfg.min_corr, has_min_corr=True,
                    jitter_mode=1, seed=int(sprm.seed))
    try:
        rep, _ = TP.check_analysis(
            O, d, cfg.n_genes, cfg.n_types, out, score_params=score_kw, min_score=cfg.min_score,
            canonical=2 if args.gene_index else 1, seed=int(nprm.seed), perm_begin=0, perm_indices=perms, ks_fast=fast,
            rounds=args.feistel_rounds or None, what=f"config {cfg.name}")
        if world > 1:
            rep["checked"].append(f"pooled null spans the slices of all {world} ranks (NCCL all-gather inside pctsea_analyze)")
        return rep
    except TP.ParityError as e:
        return {"ok": False, "failed": str(e)}


def cpu_baseline(ctx, m, ds, cfg, sprm, gene_idx, abundance, label_host, args, p_cpu=None):
    """C restatement of the reference CPU algorithm (oracle/), timed on this box's host cores: per permutation
    one Fisher-Yates shuffle + per-type two-pass weighted-KS walk, threads over types re-created per permutation
    like PCTSEA.java:2501-2519. Bounded sample: p_cpu permutations of the same ranked list (the loop is exactly
    linear in P, PCTSEA.java:1398-1429)."""
    import pctsea_parent_b200 as P
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    nprm0 = ctx.null_params(0, perm_mode=P.PERM_PHILOX, ks_mode=P.KS_EXACT, min_pass=1000)
    out = ctx.analyze(m, gene_idx, abundance, sprm, cfg.min_score, cfg.n_types, nprm0, want_scores=True, want_order=True)
    nu = out["n_pass"] - 1
    order = out["order"][:nu]
    s_r = out["score"][order]
    lab_r = label_host[order]
    # calibrate, then aim at ~15 s of CPU work
    t0 = time.perf_counter()
    O.null(s_r, lab_r, cfg.n_types, 2, perm_mode=0, seed=1, nthreads=cores)
    per = (time.perf_counter() - t0) / 2
    p_cpu = p_cpu or int(max(4, min(cfg.n_perm, 15.0 / max(per, 1e-6))))
    t0 = time.perf_counter()
    O.null(s_r, lab_r, cfg.n_types, p_cpu, perm_mode=0, seed=1, nthreads=cores)
    dt = time.perf_counter() - t0
    # stage 1 on a sample of cells (single thread, like PCTSEA.java:2305-2343)
    ns = min(cfg.n_cells, 20000)
    rp = ds.row_ptr[: ns + 1].cpu().numpy()
    nn = int(rp[-1])
    col = ds.col[:nn].cpu().numpy()
    val = ds.val[:nn].cpu().numpy()
    t0 = time.perf_counter()
    O.score(rp, col, val, cfg.n_genes, gene_idx, abundance, method=cfg.method, min_genes_cells=cfg.min_genes_cells,
            min_corr=cfg.min_corr, jitter_mode=1, seed=1, canonical=False)
    ds_ = time.perf_counter() - t0
    return {"value": nu * p_cpu / dt, "unit": "cell-permutations/s", "cores": cores, "kind": "port",
            "sample": f"{p_cpu} JDK-shuffle permutations of the same {nu}-cell ranked list, {cfg.n_types} types, "
                      f"{cores} threads over types re-created per permutation; C restatement of the reference CPU "
                      f"algorithm (not Java)",
            "null_seconds_per_perm": dt / p_cpu,
            "score_cells_per_s_1thread": ns / ds_}


def run_reference(args):
    """The reference's own CPU implementation of the path cannot run here (Java 8 + MongoDB + private SNAPSHOT
    dependencies, no JDK): this arm MEASURES the oracle port of it on the host cores, on the same workload config.
    One step = one whole analysis job, every stage timed, nothing extrapolated: scoring of all cells on one thread
    (PCTSEA.java:2309-2343), ranking, observed KS, P_step permutations (JDK Collections.shuffle + two-pass
    weighted KS per type) on 2 x cores threads re-created per permutation (PCTSEA.java:137, 2501-2519), a10-a12.
    P_step is the workload's own -perm when K such jobs fit ~7 minutes, else the largest count that fits (stated in
    `sample`). Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import oracle as O
    from pctsea_parent_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_config(args.config, 1)
    P_per_gpu = args.perm if args.perm else cfg.n_perm
    P_total = P_per_gpu * world
    cores = os.cpu_count() or 1
    threads = 2 * cores
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ds = synth.make_dataset(cfg, dev)  # data generation only; nothing of the measured path runs on the GPU
    d = ds.numpy()
    nnz = ds.nnz
    del ds
    T = cfg.n_types
    parts = {}

    def job(n_perm, seed):
        t0 = time.perf_counter()
        sc, used, nz, kept = O.score(d["row_ptr"], d["col"], d["val"], cfg.n_genes, d["gene_idx"], d["abundance"],
                                     method=cfg.method, min_genes_cells=cfg.min_genes_cells, min_corr=cfg.min_corr,
                                     jitter_mode=1, seed=1, canonical=False)
        t1 = time.perf_counter()
        order, npass = O.rank(sc, kept, cfg.min_score)
        nu = npass - 1
        s_r, lab_r = sc[order[:nu]], d["label"][order[:nu]]
        t2 = time.perf_counter()
        ores = O.ks_observed(s_r, lab_r, T)
        t3 = time.perf_counter()
        ne, _ = O.null(s_r, lab_r, T, n_perm, perm_mode=0, seed=seed, nthreads=threads)
        t4 = time.perf_counter()
        O.reduce(ores, ne)
        t5 = time.perf_counter()
        parts.update(score=t1 - t0, rank=t2 - t1, observed=t3 - t2, null=t4 - t3, reduce=t5 - t4)
        return t5 - t0, nu, npass

    # calibration = the warm-up step: a job with 4 permutations
    t_cal, nu, npass = job(4, 0)
    t_perm = parts["null"] / 4
    t_fixed = t_cal - parts["null"]
    budget_s = 420.0
    if args.steps * (t_fixed + P_total * t_perm) <= budget_s:
        P_step = P_total
    else:
        P_step = int(max(10, min(P_total, (budget_s / max(args.steps, 1) - t_fixed) / max(t_perm, 1e-9))))
    times = []
    for k in range(args.steps):
        dt, nu, npass = job(P_step, 1 + k)
        times.append(dt)
    t_step = float(np.mean(times))
    value = nu * P_step / t_step
    line = {
        "impl": "reference", "metric": "cells x permutations KS-scored per second", "value": value,
        "unit": "cell-permutations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 scores / f32 running sums", "data": "synthetic",
        "config": {"workload": workload_string(cfg, P_per_gpu, P_total), "nnz": nnz,
                   "Np": int(npass), "Nu": int(nu), "n_perm_total": P_total},
        "cpu_baseline": {"value": value, "unit": "cell-permutations/s", "cores": threads, "kind": "port",
                         "extrapolated": False, "n_perm_per_step": P_step,
                         "sample": f"each step = one whole analysis job measured on the CPU: scoring of all {cfg.n_cells} cells "
                                   f"(1 thread, as the reference), rank, observed KS, {P_step} JDK-shuffle permutations on the "
                                   f"{nu}-cell ranked list with {threads} threads (2 x {cores} cores) over types re-created per "
                                   f"permutation, a10-a12; " +
                                   ("the workload's own -perm" if P_step == P_total else
                                    f"{P_step} of the workload's {P_total} permutations so that {args.steps} steps fit {budget_s:.0f} s "
                                    f"(fewer permutations per job make the fixed stages weigh more: this understates the CPU)") +
                                   "; warm-up = one 4-permutation job; C restatement of the reference CPU algorithm, not Java"},
        "e2e": {"value": value, "unit": "cell-permutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "parts_s": dict(parts), "per_step_s": [round(t, 3) for t in times],
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", help="workload from pctsea_parent_b200.synth.CONFIGS (default B)")
    ap.add_argument("--perm", type=int, default=0, help="permutations per GPU (default: the config's)")
    ap.add_argument("--ks", default="fast", choices=["fast", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-perm", type=int, default=10_000,
                    help="permutations of the strong-scaling job timed in the same run (config C; 0 = skip)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate after the timed region (development only)")
    ap.add_argument("--feistel-rounds", type=int, default=0, help="rounds of the keyed label permutation (0 = library default)")
    ap.add_argument("--gene-index", type=int, default=1,
                    help="1 (default): build the matrix gene-major copy at load, scoring reads only the list genes; 0: CSR scan")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
