"""Seeded synthetic inputs of the BASELINE.json shapes (SURVEY.md §8 d).

HCL-shaped single-cell matrix (UMI-like counts, Zipf gene popularity, log-normal cell depth), Zipf-sized cell
types with 1 % unknown labels, a tenth of the types planted as enriched for the input list, and an input list
with integer spectral-count-like abundances. Built with torch so the large configs are generated directly in
HBM; the same code runs on the CPU at test sizes. Not part of the measured path.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import torch

MASTER_SEED = 20261019


@dataclasses.dataclass
class Config:
    name: str
    n_cells: int
    n_genes: int
    density: float
    n_types: int
    list_len: int
    method: int          # 0 PEARSONS_CORRELATION, 1 SIMPLE_SCORE
    n_perm: int
    min_genes_cells: int
    min_corr: float
    min_score: float
    config_id: int = 0


CONFIGS = {
    # BASELINE.json configs[0..3]; E reuses B's matrix with 256 lists
    "A": Config("A", 20_000, 20_000, 0.05, 40, 500, 0, 50, 4, 0.2, 0.0, 1),
    "B": Config("B", 600_000, 25_000, 0.05, 100, 1000, 0, 1000, 4, 0.2, 0.0, 2),
    "C": Config("C", 600_000, 25_000, 0.05, 100, 1000, 0, 10_000, 4, 0.2, 0.0, 2),
    "D": Config("D", 2_000_000, 30_000, 0.05, 300, 1000, 1, 10_000, 10, -1.0, 0.0, 4),
    # CPU-test sizes
    "tiny": Config("tiny", 3_000, 2_000, 0.05, 12, 120, 0, 16, 4, 0.2, 0.0, 9),
    "small": Config("small", 12_000, 4_000, 0.05, 24, 300, 0, 32, 4, 0.2, 0.0, 10),
}


@dataclasses.dataclass
class Dataset:
    cfg: Config
    row_ptr: torch.Tensor   # int64 [N+1]
    col: torch.Tensor       # int32 [nnz]
    val: torch.Tensor       # fp32 [nnz]
    label: torch.Tensor     # int32 [N], -1 = unknown
    gene_idx: torch.Tensor  # int32 [L]
    abundance: torch.Tensor  # fp32 [L]

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def numpy(self):
        return {k: getattr(self, k).cpu().numpy() for k in ("row_ptr", "col", "val", "label", "gene_idx", "abundance")}


def _gen(device, *stream) -> torch.Generator:
    g = torch.Generator(device=device)
    seed = MASTER_SEED
    for s in stream:
        seed = (seed * 1_000_003 + int(s) + 7) % (2**62)
    g.manual_seed(seed)
    return g


def make_list(cfg: Config, device="cpu", list_seed: int = 0):
    """Input list: L distinct genes, 80 % from the 5000 most popular (gene id = popularity rank),
    abundance = round(LogNormal(2, 1)) + 1."""
    g = _gen("cpu", cfg.config_id, 100, list_seed)
    top = min(5000, cfg.n_genes)
    if cfg.n_genes - top >= cfg.list_len:
        n_top = min(int(round(cfg.list_len * 0.8)), top)
        a = torch.randperm(top, generator=g)[:n_top]
        b = top + torch.randperm(cfg.n_genes - top, generator=g)[:cfg.list_len - n_top]
        genes = torch.cat([a, b])
    else:  # small test matrices: the whole gene range is "popular"
        genes = torch.randperm(cfg.n_genes, generator=g)[:cfg.list_len]
    genes = genes[torch.randperm(genes.numel(), generator=g)]
    ab = torch.round(torch.exp(2.0 + torch.randn(genes.numel(), generator=g))) + 1.0
    return genes.to(torch.int32).to(device), ab.to(torch.float32).to(device)


def make_labels(cfg: Config, device="cpu"):
    g = _gen("cpu", cfg.config_id, 200)
    T, N = cfg.n_types, cfg.n_cells
    w = 1.0 / torch.arange(1, T + 1, dtype=torch.float64)
    floor = min(50, max(2, N // (4 * T)))
    sizes = torch.clamp((w / w.sum() * (N - floor * T)).floor().long(), min=0) + floor
    sizes[0] += N - int(sizes.sum())
    lab = torch.repeat_interleave(torch.arange(T, dtype=torch.int32), sizes)
    lab = lab[torch.randperm(N, generator=g)]
    unk = torch.rand(N, generator=g) < 0.01
    lab[unk] = -1
    return lab.to(device)


def make_dataset(cfg: Config, device="cpu", chunk_cells: int = 8192, list_seed: int = 0) -> Dataset:
    dev = torch.device(device)
    N, G = cfg.n_cells, cfg.n_genes
    gene_idx, abundance = make_list(cfg, dev, list_seed)
    label = make_labels(cfg, dev)
    # gene popularity ~ Zipf(1.1) over gene id, scaled to the target mean density
    pop = 1.0 / torch.arange(1, G + 1, dtype=torch.float64) ** 1.1
    lo, hi = 0.0, 1e9
    for _ in range(80):  # scale c with mean(min(0.9, c*pop)) = density
        mid = 0.5 * (lo + hi)
        if torch.clamp(mid * pop, max=0.9).mean().item() < cfg.density:
            lo = mid
        else:
            hi = mid
    p_gene = torch.clamp(hi * pop, max=0.9).to(torch.float32).to(dev)

    n_enriched = max(1, cfg.n_types // 10)
    enriched = torch.zeros(cfg.n_types + 1, dtype=torch.bool, device=dev)
    enriched[:n_enriched * 3:3] = True  # types 0, 3, 6, ... (a mix of big and small types)
    list_cols = gene_idx.long()
    sqrt_ab = torch.sqrt(abundance)

    cols, vals, counts = [], [], []
    for c0 in range(0, N, chunk_cells):
        c1 = min(N, c0 + chunk_cells)
        n = c1 - c0
        g = _gen(dev, cfg.config_id, 300, c0)
        depth = torch.exp(0.6 * torch.randn(n, generator=g, device=dev) - 0.18)
        prob = torch.clamp(p_gene[None, :] * depth[:, None], max=0.95)
        mask = torch.rand(n, G, generator=g, device=dev) < prob
        v = torch.empty(n, G, device=dev, dtype=torch.float32).geometric_(0.45, generator=g)
        v = v * mask
        # planted signal: cells of enriched types express list genes in proportion to sqrt(abundance)
        lab_chunk = label[c0:c1].long()
        is_enr = enriched[torch.where(lab_chunk < 0, cfg.n_types, lab_chunk)]
        rows = torch.nonzero(is_enr).flatten()
        if rows.numel() > 0:
            k = 1.0 + 2.0 * torch.rand(rows.numel(), 1, generator=g, device=dev)
            noise = torch.exp(0.5 * torch.randn(rows.numel(), list_cols.numel(), generator=g, device=dev))
            planted = torch.clamp(torch.round(k * sqrt_ab[None, :] * noise), min=1.0)
            on = torch.rand(rows.numel(), list_cols.numel(), generator=g, device=dev) < 0.6
            sub = v[rows][:, list_cols]
            sub = torch.where(on, planted, sub)
            v[rows[:, None], list_cols[None, :]] = sub
        nz = torch.nonzero(v)  # row-major: columns ascending inside a row
        counts.append(torch.bincount(nz[:, 0], minlength=n))
        cols.append(nz[:, 1].to(torch.int32))
        vals.append(v[nz[:, 0], nz[:, 1]])
        del mask, v, prob, nz
    cnt = torch.cat(counts)
    row_ptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(cnt, 0)
    return Dataset(cfg, row_ptr, torch.cat(cols), torch.cat(vals), label, gene_idx, abundance)


def synthetic_ranked(n_used: int, n_types: int, seed: int = 0, simple: bool = False):
    """Ranked scores + labels without a matrix (stage-3 tests): descending scores in [0.2,1) with a few
    planted types concentrated near the top."""
    rng = np.random.default_rng(MASTER_SEED + seed)
    s = np.sort(rng.uniform(0.2, 1.0, n_used))[::-1].copy()
    if simple:
        s = s + np.sort(rng.integers(4, 60, n_used))[::-1]
    w = 1.0 / np.arange(1, n_types + 1)
    lab = rng.choice(n_types, size=n_used, p=w / w.sum()).astype(np.int32)
    top = rng.random(n_used) < np.linspace(0.5, 0.0, n_used)
    lab[top & (rng.random(n_used) < 0.5)] = 0
    lab[rng.random(n_used) < 0.01] = -1
    return s, lab
