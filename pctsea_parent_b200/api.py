"""Host-side mirror of pctsea-core's public surface for the accelerated path.

Names, argument meaning and error behaviour follow the reference classes (paths under
pctsea-core/src/main/java/edu/scripps/yates/pctsea/):
  model/InputParameters.java:13-227, model/ScoringMethod.java:8-23, model/ScoringSchema.java:14-57,
  scoring/ScoreThreshold.java, model/CellTypeClassification.java (getters), model/PCTSEAResult.java:17-64,
  PCTSEA.java:174-217 (constructors), :257-534 (run), :2692-2818 (setters).
The MongoDB repositories of the reference constructor are replaced by a SingleCellDataset: the matrix and the
cell-type labels loaded once into device-resident CSR. Everything numeric goes through the C ABI
(include/pctsea_b200.h); there is no CPU path here.
"""
from __future__ import annotations

import dataclasses
import enum
import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Context, PctseaError


class ScoringMethod(enum.Enum):
    # model/ScoringMethod.java:8-23 — (scoreName)
    SIMPLE_SCORE = "SimpleScore"
    PEARSONS_CORRELATION = "correlation"
    DOT_PRODUCT = "Dot-product"
    QUICK_SCORE = "Quick_score"
    REGRESSION = "Linear Regression"

    def getScoreName(self) -> str:
        return self.value


class ScoreThreshold:
    """scoring/ScoreThreshold.java: >= threshold when threshold >= 0, else < threshold."""

    def __init__(self, threshold: float):
        self.threshold = float(threshold)

    def getThresholdValue(self) -> float:
        return self.threshold

    def __str__(self):
        return ("<" if self.threshold < 0.0 else ">") + " " + str(self.threshold)


@dataclasses.dataclass
class ScoringSchema:
    scoringMethod: ScoringMethod
    scoringThreshold: ScoreThreshold
    minGenesCells: int

    def getScoringMethod(self):
        return self.scoringMethod

    def getScoringThreshold(self):
        return self.scoringThreshold

    def getMinGenesCells(self):
        return self.minGenesCells


class InputParameters:
    # model/InputParameters.java:14-17, :49-64
    DEFAULT_MIN_SCORE_SIMPLE_SCORE = 0.0
    DEFAULT_MIN_SCORE_PEARSON = 0.2
    DEFAULT_MIN_GENES_CELLS_SIMPLE_SCORE = 10
    DEFAULT_MIN_GENES_CELLS_PEARSON = 100
    EMAIL, OUT, PERM, EEF = "email", "out", "perm", "eef"
    MIN_SCORE, MIN_GENES_CELLS = "min_score", "min_genes_cells"
    CELL_TYPES_CLASSIFICATION, LOAD_RANDOM = "cell_types_classification", "load_random"
    PLOT_NEGATIVE_ENRICHED, DATASETS, WRITE_SCORES = "plot_negative_enriched", "datasets", "write_scores"
    UNIPROT_RELEASE, SCORING_METHOD, INPUT_DATA_TYPE = "uniprot_release", "scoring_method", "input_data_type"
    MINIMUM_CORRELATION, CREATE_ZIP_FILE = "min_corr", "create_zip"

    def __init__(self):
        self.inputDataFile: Optional[str] = None
        self.outputPrefix: Optional[str] = None
        self.loadRandom = False
        self.numPermutations = 10            # PCTSEACommandLine.java:213
        self.minCorr: Optional[float] = -1.0  # PCTSEACommandLine.java:273
        self.scoringSchemas: List[ScoringSchema] = []
        # new, optional knobs (defaults keep the reference's behaviour; SURVEY.md §5 "Config / flags")
        self.permutationMode = "PHILOX"      # or "REPLAY"
        self.ksMode = "FAST"                 # or "EXACT"
        self.seed = 0

    def addScoringSchema(self, scoringMethod: ScoringMethod, minScore: float, minGenesCells: int):
        self.scoringSchemas.append(ScoringSchema(scoringMethod, ScoreThreshold(minScore), minGenesCells))

    def setInputDataFile(self, f):
        self.inputDataFile = f

    def setNumPermutations(self, n):
        self.numPermutations = int(n)

    def setMinCorr(self, v):
        self.minCorr = v

    def setOutputPrefix(self, p):
        self.outputPrefix = p

    def setLoadRandom(self, b):
        self.loadRandom = bool(b)


def read_experimental_expressions_file(path: str):
    """GeneExpressionsRetriever.readExperimentalExpressionsFile (GeneExpressionsRetriever.java:358-409):
    skip blank lines, skip the first non-blank line, TAB-separated, gene = col0 upper-cased,
    abundance = Float.valueOf(col1); ids in file order."""
    genes, ab = [], []
    num_line = 0
    with open(path, "r") as fh:
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if line.strip() == "":
                continue
            num_line += 1
            if num_line == 1:
                continue
            if "\t" not in line:
                raise IOError(f"Error reading input file at line {num_line}: At least 2 columns separated by TAB are needed")
            sp = line.split("\t")
            if len(sp) < 2:
                raise IOError(f"Error reading input file at line {num_line}: At least 2 columns separated by TAB are needed")
            genes.append(sp[0].upper())
            ab.append(np.float32(sp[1]))
    return genes, np.asarray(ab, dtype=np.float32)


def _java_float_str(v) -> str:
    """What `"" + (double) floatValue` prints closely enough for Float.valueOf to read it back exactly."""
    v = float(v)
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    return repr(v)


def _java_number_str(x) -> str:
    """Float.toString / Double.toString of a numpy float32 / float64: shortest digits that round-trip, plain decimal
    for 1e-3 <= |x| < 1e7 and d.dddE±n otherwise, always at least one fractional digit."""
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    ax = abs(float(x))
    if 1e-3 <= ax < 1e7:
        r = np.format_float_positional(x, unique=True, trim="0")     # trim="0": keep one digit after the point
        return r
    r = np.format_float_scientific(x, unique=True, trim="0")          # e.g. 1.e-05 / 1.5e+300 with trim rules
    mant, ex = r.split("e")
    if mant.endswith("."):
        mant += "0"
    return f"{mant}E{int(ex)}"


def _java_double_str(v) -> str:
    return _java_number_str(np.float64(v))


def _java_float32_str(v) -> str:
    return _java_number_str(np.float32(v))


def get_output_txt_file_name(file_name: str, prefix: str, scoring_method: "ScoringMethod") -> str:
    """PCTSEAUtils.getOutputTXTFile (PCTSEAUtils.java:109-114)."""
    return f"{prefix}_{scoring_method.getScoreName()}_{file_name}.txt"


def write_score_calculation_file(path: str, cell_type_name: str, a: np.ndarray, b: np.ndarray, supremum_x: int,
                                 secondary_supremum_x: int = 0):
    """EnrichmentWeigthedScoreParallel.writeScoreCalculationFile (:355-410): the two cumulative series with runs of
    equal y collapsed to their end points, then the supremum line(s). x = ranked position, 1-based."""
    with open(path, "w", encoding="utf-8") as f:
        f.write("-\tcell #\tCumulative Probability [Fn(X)]\n")
        for label, series in ((cell_type_name, a), ("others", b)):
            series = np.asarray(series, np.float32)
            bits = series.view(np.uint32)
            prev = -1
            for i in range(series.size):
                # Float.compare(y, previous.y) == 0: same bits (NaNs are canonical in the reference; -0.0 != 0.0)
                if prev >= 0 and bits[i] == bits[prev]:
                    prev = i
                    continue
                if prev >= 0:
                    f.write(f"{label}\t{prev + 1}\t{_java_float32_str(series[prev])}\n")
                f.write(f"{label}\t{i + 1}\t{_java_float32_str(series[i])}\n")
                prev = i
        if supremum_x >= 1:
            f.write(f"supremum\t{supremum_x}\t{_java_float32_str(a[supremum_x - 1])}\n")
            f.write(f"supremum\t{supremum_x}\t{_java_float32_str(b[supremum_x - 1])}\n")
        if secondary_supremum_x >= 1:
            f.write(f"secondary supremum\t{secondary_supremum_x}\t{_java_float32_str(a[secondary_supremum_x - 1])}\n")
            f.write(f"secondary supremum\t{secondary_supremum_x}\t{_java_float32_str(b[secondary_supremum_x - 1])}\n")


def write_score_distribution_file(path: str, score_name: str, scores_from_cell_type: Sequence[float]):
    """writeScoreDistributionFile (:437-447): the type's non-NaN ranking scores, in ranked order."""
    with open(path, "w") as f:
        f.write(score_name + "\n")
        for sc in scores_from_cell_type:
            f.write(_java_double_str(sc) + "\n")


def write_num_genes_histogram_file(path: str, score_name: str, num_genes: Sequence[int]):
    """writeNumGenesHistogramFile (:412-435): histogram of the number of genes used for the score over the type's
    cells, plain and cumulative from the top. (The reference iterates a trove hash map, so its row order is the map's;
    here rows are in ascending key order.)"""
    keys, counts = np.unique(np.asarray(num_genes, np.int64), return_counts=True)
    with open(path, "w") as f:
        f.write("-\t# of genes with " + score_name + " > threshold\t# cells\n")
        for k, c in zip(keys, counts):
            f.write(f"# genes\t{int(k)}\t{int(c)}\n")
        acc = np.cumsum(counts[::-1])[::-1]
        for k, c in zip(keys, acc):
            f.write(f"# genes or more\t{int(k)}\t{int(c)}\n")


def get_random_scores_file_name(prefix: str, scoring_method: "ScoringMethod", max_iterations: int) -> str:
    """PCTSEA.getRandomScoresFile (PCTSEA.java:1911-1914), without the time-stamp folder."""
    return f"{prefix}_{scoring_method.getScoreName()}_{max_iterations}_random_scores.txt"


def print_to_random_distribution_file(path: str, cell_types: Sequence["CellTypeClassification"]):
    """PCTSEA.printToRandomDistributionFile (PCTSEA.java:1877-1909): one row per type, sorted by name:
    type, real score, diff, mean, std, then (score, dStat) per permutation."""
    with open(path, "w") as fh:
        fh.write("cell type\treal score\tdiff\tavg random score\tstd random score\trandom scores\n")
        for ct in sorted(cell_types, key=lambda c: c.getName()):
            real = np.float32(ct.getEnrichmentScore())
            rnd = np.asarray(ct.randomEnrichmentScores, dtype=np.float32)
            dst = np.asarray(ct.randomKSTestStatistics, dtype=np.float32)
            mean = np.float32(0)
            for v in rnd:                      # Maths.mean(TFloatList): float running sum / size
                mean = np.float32(mean + v)
            with np.errstate(invalid="ignore", divide="ignore"):
                mean = np.float32(mean / np.float32(rnd.size)) if rnd.size else np.float32("nan")
                std = float(np.std(rnd.astype(np.float64), ddof=1)) if rnd.size > 1 else float("nan")
            diff = float(real) - float(mean)
            fh.write(f"{ct.getName()}\t{_java_float_str(real)}\t{_java_float_str(diff)}\t{_java_float_str(mean)}\t{_java_float_str(std)}")
            for sc, d in zip(rnd, dst):
                fh.write(f"\t{_java_float_str(sc)}\t{_java_float_str(d)}")
            fh.write("\n")


def read_random_distribution_file(path: str):
    """PCTSEA.readRandomDistributionFile (PCTSEA.java:1834-1869): skip the header; columns 5.. are
    (score, dStat) pairs parsed as float. Returns {type name: (scores fp32[P], dstats fp32[P])}."""
    out = {}
    with open(path, "r") as fh:
        for num, line in enumerate(fh, start=1):
            if num == 1:
                continue
            sp = line.rstrip("\n").split("\t")
            if len(sp) < 5:
                continue
            vals = [np.float32(x) for x in sp[5:]]
            out[sp[0]] = (np.asarray(vals[0::2], np.float32), np.asarray(vals[1::2], np.float32))
    return out


def p_adjust_bh(pvalues: Sequence[float]) -> np.ndarray:
    """Benjamini-Hochberg adjusted values, the way R's p.adjust(p, "BH") computes them: NaNs stay NaN and do not count
    towards n; sorted descending, cumulative minimum of n/i * p, capped at 1. PCTSEA.java:1551-1557 calls
    edu.scripps.yates.utilities.maths.PValueCorrection.pAdjust(.., BH) — an un-vendored dependency of the reference
    (utilities artifact, no source under /root/reference) — restated here from the published algorithm."""
    p = np.asarray(pvalues, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    n = int(ok.sum())
    if n == 0:
        return out
    idx = np.nonzero(ok)[0]
    order = idx[np.argsort(-p[idx], kind="stable")]           # descending
    ranks = np.arange(n, 0, -1, dtype=np.float64)              # n, n-1, ..., 1
    adj = np.minimum(1.0, np.minimum.accumulate(n / ranks * p[order]))
    out[order] = adj
    return out


def hypergeometric_upper_cumulative(N: int, K: int, n: int, k: int) -> float:
    """commons-math3 HypergeometricDistribution(N, K, n).upperCumulativeProbability(k) = P(X >= k), clamped at 0 like
    PCTSEA.java:2221-2225. O(T) host-side work on the counts the GPU returns (pctsea_last_type_counts)."""
    from scipy.stats import hypergeom
    p = float(hypergeom.sf(k - 1, N, K, n))
    return max(p, 0.0)


class SingleCellDataset:
    """The dataset the reference keeps in MongoDB (db/Expression.java, db/SingleCell.java), held as
    device-resident CSR: rows = cells in DB order, columns = upper-cased gene names."""

    def __init__(self, ctx: Context, row_ptr, col, val, gene_names: Sequence[str], cell_types: Sequence[str],
                 cell_names: Optional[Sequence[str]] = None, gene_major: bool = False):
        self.ctx = ctx
        self.gene_names = [g.upper() for g in gene_names]
        self.gene_index: Dict[str, int] = {g: i for i, g in enumerate(self.gene_names)}
        self.cell_names = list(cell_names) if cell_names is not None else None
        # model/CellTypes.getCellTypeID: string -> int id registry; empty / None = unknown (-1)
        self.type_names: List[str] = sorted({t for t in cell_types if t})
        tid = {t: i for i, t in enumerate(self.type_names)}
        self.label = np.asarray([tid.get(t, -1) if t else -1 for t in cell_types], dtype=np.int32)
        self.matrix = ctx.load_csr(row_ptr, col, val, len(self.gene_names))
        self.matrix.set_labels(self.label, len(self.type_names))
        if gene_major:   # once per dataset: the scoring stage then reads only the list's genes (score_gm.cu)
            self.matrix.build_gene_index()

    @classmethod
    def from_arrays(cls, ctx: Context, row_ptr, col, val, n_genes: int, label: np.ndarray, n_types: int,
                    gene_major: bool = False):
        self = cls.__new__(cls)
        self.ctx = ctx
        self.gene_names = [f"G{i}" for i in range(n_genes)]
        self.gene_index = None
        self.cell_names = None
        self.type_names = [f"type{i}" for i in range(n_types)]
        self.label = np.asarray(label, dtype=np.int32)
        self.matrix = ctx.load_csr(row_ptr, col, val, n_genes)
        self.matrix.set_labels(self.label, n_types)
        if gene_major:
            self.matrix.build_gene_index()
        return self

    def map_list(self, genes: Sequence[str]) -> np.ndarray:
        if self.gene_index is None:
            return np.asarray([int(g[1:]) if g.startswith("G") else -1 for g in genes], dtype=np.int32)
        return np.asarray([self.gene_index.get(g.upper(), -1) for g in genes], dtype=np.int32)


class CellTypeClassification:
    """model/CellTypeClassification.java — the fields the hot path sets, with the reference's getter names."""

    def __init__(self, name: str, cellTypeID: int, r):
        self.name, self.cellTypeID = name, cellTypeID
        self.enrichmentScore = float(r["ews"])
        self.dStatistic = float(r["dstat"])
        self.supremumX = int(r["supX"])
        self.normalizedSupremumX = float(r["norm_supX"])
        self.sizeA, self.sizeB = int(r["sizeA"]), int(r["sizeB"])
        self.normalizedEnrichmentScore = float(r["norm_ews"])
        self.enrichmentSignificance = float(r["p_emp"])
        self.enrichmentFDR = float(r["fdr"])
        # observed-only extras: setSecondaryEnrichment, setKSTestPvalue (EnrichmentWeigthedScoreParallel.java:228,310),
        # setEnrichmentUnweightedScore (PCTSEA.java:2571-2621: 0 for an empty type, NaN for sizes <= 1)
        self.secondaryEnrichmentScore = float(r["sec_ews"])
        self.secondarySupremumX = int(r["sec_supX"])
        self.ksTestPvalue = float(r["ks_d"])
        self.enrichmentUnweightedScore = 0.0 if self.sizeA == 0 else float(np.float32(r["ks_d"]))
        # BH-corrected KS value and its stars (PCTSEA.java:1551-1570), filled by PCTSEA.run over the round's types
        self.ksTestCorrectedPvalue = float("nan")
        self.ksTestSignificancyString = ""
        self.randomEnrichmentScores: Optional[np.ndarray] = None
        self.randomKSTestStatistics: Optional[np.ndarray] = None
        # calculateHyperGeometricStatistics (PCTSEA.java:2169-2264), filled by PCTSEA.run from the GPU's counts
        self.numCellsOfType = 0
        self.numCellsOfTypePassingCorrelationThreshold = 0
        self.hypergeometricPValue = float("nan")
        self.casimirsEnrichmentScore = float("nan")

    def getName(self): return self.name
    def getCellTypeID(self): return self.cellTypeID
    def getEnrichmentScore(self): return self.enrichmentScore
    def getKSTestDStatistic(self): return self.dStatistic
    def getSupremumX(self): return self.supremumX
    def getNormalizedSupremumX(self): return self.normalizedSupremumX
    def getSizeA(self): return self.sizeA
    def getSizeB(self): return self.sizeB
    def getNormalizedEnrichmentScore(self): return self.normalizedEnrichmentScore
    def getEnrichmentSignificance(self): return self.enrichmentSignificance
    def getEnrichmentFDR(self): return self.enrichmentFDR
    def getRandomEnrichmentScores(self): return self.randomEnrichmentScores
    def getNumCellsOfType(self): return self.numCellsOfType
    def getNumCellsOfTypePassingCorrelationThreshold(self): return self.numCellsOfTypePassingCorrelationThreshold
    def getHypergeometricPValue(self): return self.hypergeometricPValue
    def getCasimirsEnrichmentScore(self): return self.casimirsEnrichmentScore
    def getSecondaryEnrichmentScore(self): return self.secondaryEnrichmentScore
    def getSecondarySupremumX(self): return self.secondarySupremumX
    def getKSTestPvalue(self): return self.ksTestPvalue
    def getKSTestCorrectedPvalue(self): return self.ksTestCorrectedPvalue
    def getKSTestSignificancyString(self): return self.ksTestSignificancyString
    def getEnrichmentUnweightedScore(self): return self.enrichmentUnweightedScore


class PCTSEAResult:
    def __init__(self):
        self.significantTypes: List[CellTypeClassification] = []
        self.cellTypesPerRound: List[List[CellTypeClassification]] = []
        self.timings: List[dict] = []

    def getSignificantTypes(self):
        return self.significantTypes


class PCTSEA:
    MIN_CELLS_PASSING_SCORE_TRESHOLD = 1000       # PCTSEA.java:138
    CELL_TYPE_FDR_SIGNIFICANCE_THRESHOLD = 0.05   # PCTSEA.java:167

    def __init__(self, inputParameters: Optional[InputParameters], dataset: SingleCellDataset):
        self.dataset = dataset
        self.ctx = dataset.ctx
        self.sequentialScoringSchemas: List[ScoringSchema] = []
        self.maxIterations = 10
        self.minCorr: Optional[float] = -1.0
        self.experimentExpressionFile: Optional[str] = None
        self.statusListener: Optional[Callable[[str], None]] = None
        self.permutationMode, self.ksMode, self.seed = "PHILOX", "FAST", 0
        self.inputGenes: Optional[Sequence[str]] = None
        self.inputAbundance: Optional[np.ndarray] = None
        self.keepNull = False
        self.loadRandomDistributionsIfExist = False
        self.prefix: Optional[str] = None
        self.outputFolder = "."
        # per-type curve / distribution files of EnrichmentWeigthedScoreParallel.java:325-349 (off by default: they are
        # reporting output; needs a prefix). plotNegativeEnrichedCellTypes as in the reference's constructor flag.
        self.writeCellTypeFiles = False
        self.plotNegativeEnrichedCellTypes = False
        if inputParameters is not None:
            self.sequentialScoringSchemas = list(inputParameters.scoringSchemas)
            self.experimentExpressionFile = inputParameters.inputDataFile
            self.maxIterations = inputParameters.numPermutations
            self.minCorr = inputParameters.minCorr
            self.permutationMode = inputParameters.permutationMode
            self.ksMode = inputParameters.ksMode
            self.seed = inputParameters.seed
            self.loadRandomDistributionsIfExist = inputParameters.loadRandom
            self.prefix = inputParameters.outputPrefix

    # setters, PCTSEA.java:2692-2818
    def setExperimentExpressionFile(self, f): self.experimentExpressionFile = f
    def addScoreSchema(self, s: ScoringSchema): self.sequentialScoringSchemas.append(s)
    def setMaxIterations(self, n): self.maxIterations = int(n)
    def setMinimumCorrelation(self, v): self.minCorr = v
    def setLoadRandomDistributionsIfExist(self, b): self.loadRandomDistributionsIfExist = bool(b)
    def setPrefix(self, p): self.prefix = p
    def setStatusListener(self, cb): self.statusListener = cb
    def setInputList(self, genes: Sequence[str], abundance): self.inputGenes, self.inputAbundance = genes, np.asarray(abundance, np.float32)

    def logStatus(self, msg: str):
        if self.statusListener:
            self.statusListener(msg)

    def _write_cell_type_files(self, types, out, method):
        """The three per-type text files of EnrichmentWeigthedScoreParallel.java:325-349 for every type with an
        enrichment score (nk > 1) — skipping negatively enriched types unless plotNegativeEnrichedCellTypes — from
        the curves still resident on the GPU (pctsea_last_observed_curve) and the ranked scores."""
        import os
        n_used = int(out["n_pass"]) - 1            # the ranked list drops its last cell (PCTSEA.java:421-424)
        order = out["order"][:n_used]
        lab_r = self.dataset.label[order]
        sc_r = out["score"][order]
        used_r = out["n_genes_used"][order]
        name = method.getScoreName()
        for ct in types:
            if math.isnan(ct.enrichmentScore):
                continue
            if not self.plotNegativeEnrichedCellTypes and ct.enrichmentScore < 0:
                continue
            a, b = self.ctx.last_observed_curve(ct.cellTypeID, n_used)
            base = lambda suffix: os.path.join(self.outputFolder, get_output_txt_file_name(ct.name + suffix, self.prefix, method))
            write_score_calculation_file(base("_ews"), ct.name, a, b, ct.supremumX, ct.secondarySupremumX)
            mine = lab_r == ct.cellTypeID
            s_mine = sc_r[mine]
            write_score_distribution_file(base("_corr"), name, s_mine[~np.isnan(s_mine)])
            write_num_genes_histogram_file(base("_genes_per_cell_hist"), name, used_r[mine])

    def run(self) -> PCTSEAResult:
        """PCTSEA.run (PCTSEA.java:257-534), the per-schema loop :348-475 with the bodies of a2-a12 on the GPU."""
        if self.inputGenes is None:
            if self.experimentExpressionFile is None:
                raise ValueError("no input list")
            self.inputGenes, self.inputAbundance = read_experimental_expressions_file(self.experimentExpressionFile)
        gene_idx = self.dataset.map_list(self.inputGenes)
        result = PCTSEAResult()
        if self.minCorr is None:
            # AnalyzeView.java:659-662 can hand a null Double; PCTSEA.java:379-381 unboxes it
            raise TypeError("minCorr is null (NullPointerException on unboxing in the reference)")
        for schema in self.sequentialScoringSchemas:
            method = schema.getScoringMethod()
            if method not in (ScoringMethod.PEARSONS_CORRELATION, ScoringMethod.SIMPLE_SCORE):
                raise ValueError("Method " + method.getScoreName() + " still not supported.")
            self.logStatus("Calculating " + method.getScoreName() + "...")
            sprm = _lib.ScoreParams(
                _lib.METHOD_PEARSONS_CORRELATION if method == ScoringMethod.PEARSONS_CORRELATION else _lib.METHOD_SIMPLE_SCORE,
                schema.getMinGenesCells(), float(self.minCorr), 1, _lib.JITTER_PHILOX, self.seed)
            nprm = Context.null_params(
                self.maxIterations,
                perm_mode=_lib.PERM_PHILOX if self.permutationMode == "PHILOX" else _lib.PERM_REPLAY,
                ks_mode=_lib.KS_FAST if self.ksMode == "FAST" else _lib.KS_EXACT,
                seed=self.seed, min_pass=self.MIN_CELLS_PASSING_SCORE_TRESHOLD)
            import os
            rnd_file = None
            if self.prefix is not None:
                rnd_file = os.path.join(self.outputFolder, get_random_scores_file_name(self.prefix, method, self.maxIterations))
            loaded = None
            if self.loadRandomDistributionsIfExist and rnd_file and os.path.exists(rnd_file) and os.path.getsize(rnd_file) > 0:
                loaded = read_random_distribution_file(rnd_file)     # PCTSEA.java:1385-1387
            try:
                if loaded is None:
                    want_files = self.writeCellTypeFiles and self.prefix is not None
                    out = self.ctx.analyze(self.dataset.matrix, gene_idx, self.inputAbundance, sprm,
                                           schema.getScoringThreshold().getThresholdValue(),
                                           self.dataset.matrix.n_types, nprm,
                                           want_null=self.keepNull or rnd_file is not None,
                                           want_scores=want_files, want_order=want_files)
                else:
                    # observed scores on the GPU, null from the file, a10-a12 on the GPU (pctsea_reduce_null)
                    nprm0 = Context.null_params(0, min_pass=self.MIN_CELLS_PASSING_SCORE_TRESHOLD)
                    out = self.ctx.analyze(self.dataset.matrix, gene_idx, self.inputAbundance, sprm,
                                           schema.getScoringThreshold().getThresholdValue(),
                                           self.dataset.matrix.n_types, nprm0)
                    T = self.dataset.matrix.n_types
                    Pn = max((v[0].size for v in loaded.values()), default=0)
                    null = np.full((T, Pn), np.nan, np.float32)
                    nullD = np.full((T, Pn), np.nan, np.float32)
                    for t, name in enumerate(self.dataset.type_names):
                        if name in loaded and loaded[name][0].size == Pn:
                            null[t], nullD[t] = loaded[name]
                    out["results"] = self.ctx.reduce_null(out["results"], null, 0)
                    out["null"], out["nullD"] = null, nullD
            except PctseaError as e:
                if e.code in (_lib.PCTSEA_ETOOFEW, _lib.PCTSEA_EMINGENES):
                    raise ValueError(e.msg) from e  # IllegalArgumentException in the reference
                raise
            result.timings.append(self.ctx.timings())
            types = []
            n_of_type, n_of_type_pass, n_kept = self.ctx.last_type_counts(self.dataset.matrix.n_types)
            K = int(out["n_pass"])
            for t, name in enumerate(self.dataset.type_names):
                if n_of_type[t] == 0:
                    continue   # no CellTypeClassification for a type without cells after the filter (PCTSEA.java:2172-2174)
                ct = CellTypeClassification(name, t, out["results"][t])
                ct.numCellsOfType = int(n_of_type[t])
                ct.numCellsOfTypePassingCorrelationThreshold = int(n_of_type_pass[t])
                ct.hypergeometricPValue = hypergeometric_upper_cumulative(n_kept, K, ct.numCellsOfType,
                                                                           ct.numCellsOfTypePassingCorrelationThreshold)
                if ct.numCellsOfType > 0 and K > 0:
                    x = (1.0 * ct.numCellsOfTypePassingCorrelationThreshold / K) / (1.0 * ct.numCellsOfType / n_kept)
                    ct.casimirsEnrichmentScore = float(np.float32(math.log(x) / math.log(2))) if x > 0 else float("-inf")
                if "null" in out:
                    ct.randomEnrichmentScores = out["null"][t]
                    ct.randomKSTestStatistics = out["nullD"][t]
                types.append(ct)
            if rnd_file is not None and loaded is None:
                print_to_random_distribution_file(rnd_file, types)   # PCTSEA.java:1431
            if self.writeCellTypeFiles and self.prefix is not None and "order" in out:
                self._write_cell_type_files(types, out, method)
            # BH correction of the KS values over the round's cell types (PCTSEA.java:1551-1570)
            for ct, padj in zip(types, p_adjust_bh([c.ksTestPvalue for c in types])):
                ct.ksTestCorrectedPvalue = float(padj)
                ct.ksTestSignificancyString = "***" if padj < 0.001 else "**" if padj < 0.01 else "*" if padj < 0.05 else ""
            # PCTSEA.java:2476-2482: sorted by enrichment score, descending (Double.compare: NaN first)
            types.sort(key=lambda c: (0 if math.isnan(c.enrichmentScore) else 1, -c.enrichmentScore if not math.isnan(c.enrichmentScore) else 0.0))
            result.cellTypesPerRound.append(types)
        # getIntersection (PCTSEA.java:574-604): significant in every round, FDR <= 0.05
        sig = None
        for types in result.cellTypesPerRound:
            names = {c.name for c in types if not math.isnan(c.enrichmentFDR) and c.enrichmentFDR <= self.CELL_TYPE_FDR_SIGNIFICANCE_THRESHOLD}
            sig = names if sig is None else (sig & names)
        last = result.cellTypesPerRound[-1] if result.cellTypesPerRound else []
        result.significantTypes = [c for c in last if sig and c.name in sig]
        return result
