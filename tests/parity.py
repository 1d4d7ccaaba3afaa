"""The parity gate: one whole analysis from the CUDA library compared with the CPU oracle, at any size.

Test infrastructure (it executes oracle/): used by the `-m gpu` tests at BASELINE.json's full sizes and by bench.py
after its timed region (no number is reported before the gate passes). Nothing in the product
package imports this module.

What is compared, and how (north_star's bars):
  a1-a3  scores of ALL cells, bit for bit, against the oracle's canonical form of the scoring kernel that ran
         (canonical=2 gene-major / 1 CSR scan), n_genes_used and kept exactly; and <= 1e-9 relative against the
         reference-faithful sequential SimpleRegression restatement on a sample of cells
  a4-a7  n_pass and the whole ranking, exactly
  a8     every observed field of every cell type (ews, dstat, supremum index, secondary supremum, KS D ...), bit for bit
  a9     the null of the permutation indices asked for, bit for bit against the oracle's restatement of the same
         (permutation source, arithmetic) pair: PHILOX+FAST, PHILOX+EXACT or REPLAY+EXACT
  a10-12 p-values, normalised scores and FDR recomputed by the oracle from the GPU's own null, bit for bit
"""
from __future__ import annotations

import time

import numpy as np

from tests.helpers import OBSERVED_FIELDS, bits32, bits64


class ParityError(AssertionError):
    pass


def _same_f64(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(bits64(a)[~na], bits64(b)[~nb]))


def _same_f32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return bool(a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(bits32(a)[~na], bits32(b)[~nb]))


def _same_field(a, b):
    if a.dtype == np.float32:
        return _same_f32(a, b)
    if a.dtype == np.float64:
        return _same_f64(a, b)
    return bool(np.array_equal(a, b))


def oracle_fast_null_column(O, s_r, lab_r, T, seed, p, rounds=None, frac_bits=None):
    """The oracle's PHILOX+FAST null of global permutation index p: [T] fp32."""
    kw = {} if not rounds else {"rounds": rounds}
    F = O.fast_frac_bits(s_r) if frac_bits is None else frac_bits
    return O.ks_null_fast_one(s_r, O.perm_labels(seed, p, lab_r, T, **kw), T, F)


def check_analysis(O, d, cfg_n_genes, n_types, out, *, score_params, min_score, canonical, seed, perm_begin=0,
                   perm_indices=(), ks_fast=True, rounds=None, mean_policy=0, sequential_sample=2000,
                   what="analysis"):
    """d: host CSR + label + list (numpy dict: row_ptr, col, val, label, gene_idx, abundance); out: the dict
    Context.analyze returned with want_null, want_scores and want_order; score_params = dict(method=, min_genes_cells=,
    min_corr=, has_min_corr=, jitter_mode=, seed=). perm_indices are GLOBAL permutation indices inside
    [perm_begin, perm_begin + out['null'].shape[1]). Returns a report dict; raises ParityError on any mismatch."""
    rep = {"what": what, "checked": [], "ok": True}
    t_begin = time.perf_counter()

    def fail(msg):
        rep["ok"] = False
        rep["failed"] = msg
        raise ParityError(f"{what}: {msg}")

    # ---- stage 1
    sc, used, nz, kept = O.score(d["row_ptr"], d["col"], d["val"], cfg_n_genes, d["gene_idx"], d["abundance"],
                                 canonical=canonical, **score_params)
    if not _same_f64(out["score"], sc):
        bad = np.flatnonzero(~((np.isnan(sc) & np.isnan(out["score"])) | (bits64(sc) == bits64(out["score"]))))
        fail(f"scores differ from the oracle in {bad.size} of {sc.size} cells (first {bad[:5]})")
    if not (np.array_equal(out["n_genes_used"], used) and np.array_equal(out["kept"], kept)):
        fail("n_genes_used / kept differ from the oracle")
    rep["n_cells"] = int(sc.size)
    rep["checked"].append("scores(all cells, bit-exact)")
    if sequential_sample:
        # the reference-faithful sequential update on a sample of cells: <= 1e-9 relative on regular cells
        rng = np.random.default_rng(1)
        cells = np.sort(rng.choice(sc.size, size=min(sequential_sample, sc.size), replace=False))
        rp = d["row_ptr"]
        sub_ptr = np.zeros(cells.size + 1, np.int64)
        sub_ptr[1:] = np.cumsum(rp[cells + 1] - rp[cells])
        idx = np.concatenate([np.arange(rp[c], rp[c + 1]) for c in cells]) if cells.size else np.zeros(0, np.int64)
        seq = O.score(sub_ptr, d["col"][idx], d["val"][idx], cfg_n_genes, d["gene_idx"], d["abundance"],
                      canonical=False, **{**score_params, "jitter_mode": 0})[0]
        ok = ~np.isnan(seq) & ~np.isnan(sc[cells])
        if ok.any():
            rel = float(np.max(np.abs(seq[ok] - sc[cells][ok]) / np.maximum(1.0, np.abs(seq[ok]))))
            rep["max_rel_vs_sequential"] = rel
            if rel > 1e-9:
                fail(f"scores differ from the sequential SimpleRegression restatement by {rel:.3g} relative (> 1e-9)")
        rep["checked"].append(f"scores({int(ok.sum())} sampled cells vs sequential, <= 1e-9 rel)")

    # ---- stage 2
    order, npass = O.rank(sc, kept, min_score)
    if npass != out["n_pass"] or not np.array_equal(order, out["order"]):
        fail(f"ranking differs from the oracle (n_pass {out['n_pass']} vs {npass})")
    nu = npass - 1
    rep["n_pass"], rep["n_used"] = int(npass), int(nu)
    rep["checked"].append("ranking(exact)")

    # ---- stage 3, observed
    s_r, lab_r = sc[order[:nu]], np.asarray(d["label"])[order[:nu]]
    ores = O.ks_observed(s_r, lab_r, n_types)
    for f in OBSERVED_FIELDS:
        if not _same_field(out["results"][f], ores[f]):
            fail(f"observed field {f} differs from the oracle")
    rep["checked"].append("observed(all types, all fields, bit-exact)")

    # ---- stage 3, null columns
    if len(perm_indices):
        null = out["null"]
        F = O.fast_frac_bits(s_r) if ks_fast else None
        for p in perm_indices:
            col = null[:, p - perm_begin]
            if ks_fast:
                exp = oracle_fast_null_column(O, s_r, lab_r, n_types, seed, p, rounds, F)
            else:
                kw = {} if not rounds else {"rounds": rounds}
                exp = O.ks_null_one(s_r, O.perm_labels(seed, p, lab_r, n_types, **kw), n_types)[0]
            if not _same_f32(col, exp):
                fail(f"null of permutation {p} differs from the oracle")
        rep["null_perms_checked"] = [int(p) for p in perm_indices]
        rep["checked"].append(f"null({len(perm_indices)} permutations, bit-exact)")

    # ---- a10-a12 from the GPU's own null
    if "null" in out and out["null"] is not None and perm_begin == 0:
        red = O.reduce(ores, out["null"], mean_policy=mean_policy)
        for f in ("p_emp", "norm_ews", "fdr"):
            if not _same_field(out["results"][f], red[f]):
                fail(f"{f} differs from the oracle's reduction of the same null")
        rep["checked"].append("p/norm/FDR(exact given the null)")
    rep["seconds"] = time.perf_counter() - t_begin
    return rep, (s_r, lab_r, ores)
