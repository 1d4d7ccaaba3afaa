"""pctsea_parent_b200 — B200 (sm_100a) implementation of PCTSEA's hot path behind pctsea-core's API.

The numeric path is libpctsea_b200.so (CUDA, C ABI in include/pctsea_b200.h). Importing the package does not
need a GPU; creating a Context does, and fails loudly without one (no CPU fallback).
"""
from ._lib import (Context, Matrix, PctseaError, ScoreParams, NullParams, TYPE_RESULT_DTYPE, load_library,  # noqa: F401
                   METHOD_PEARSONS_CORRELATION, METHOD_SIMPLE_SCORE, JITTER_NAN, JITTER_PHILOX, PERM_REPLAY,
                   PERM_PHILOX, KS_EXACT, KS_FAST, OPT_OBS_OVERLAP, OPT_SCORE_PATH, OPT_GM_UNROLL, OPT_SCORE_WARPS,
                   OPT_PERM_CHUNK, OPT_NULL_THREADS, OPT_NULL_CTAS_PER_SM, OPT_NULL_MIN_THREADS, OPT_DEBUG_GAPS, OPT_SHARD_SCORE, OPT_OBS_CHAIN_MB,
                   OPT_NULL_FINE_ENTRIES, OPT_SHARD_OBSERVED)
from .api import (PCTSEA, InputParameters, ScoringMethod, ScoringSchema, ScoreThreshold, SingleCellDataset,  # noqa: F401
                  CellTypeClassification, PCTSEAResult, read_experimental_expressions_file)

__all__ = ["Context", "Matrix", "PctseaError", "PCTSEA", "InputParameters", "ScoringMethod", "ScoringSchema",
           "ScoreThreshold", "SingleCellDataset", "CellTypeClassification", "PCTSEAResult"]
