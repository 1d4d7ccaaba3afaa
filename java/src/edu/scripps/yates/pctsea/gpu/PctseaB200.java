package edu.scripps.yates.pctsea.gpu;

/**
 * JNI face of libpctsea_b200.so (include/pctsea_b200.h), Java 8 (pctsea-core/pom.xml:13-15). One instance = one
 * pctsea_ctx = one GPU; a ctx is not re-entrant, so an instance belongs to one PCTSEA object (or is locked by its
 * owner). Every native method maps 1:1 to the C entry point named beside it and is implemented in
 * java/jni/pctsea_b200_jni.c. Error codes become exceptions in the shim: PCTSEA_EINVAL, PCTSEA_ETOOFEW and
 * PCTSEA_EMINGENES -> IllegalArgumentException (the reference throws the same type with the same messages,
 * PCTSEA.java:389-395, model/SingleCell.java:347-352), everything else -> IllegalStateException. There is no CPU
 * fallback: without a CUDA device the constructor throws.
 *
 * Arrays are plain Java arrays (pinned with GetPrimitiveArrayCritical for the duration of the call); matrices above
 * 2^31 - 1 non-zeros (BASELINE config D) are loaded from direct ByteBuffers (loadCsrDirect).
 */
public final class PctseaB200 implements AutoCloseable {
	static {
		System.loadLibrary("pctsea_b200_jni"); // links against libpctsea_b200.so
	}

	/** field order of one row of the result matrix filled by analyze / enrich / reduceNull (pctsea_type_result) */
	public static final int EWS = 0, DSTAT = 1, SUP_X = 2, NORM_SUP_X = 3, SIZE_A = 4, SIZE_B = 5, RAW_SUP = 6,
			NORM_EWS = 7, P_EMP = 8, FDR = 9, SEC_EWS = 10, SEC_SUP_X = 11, KS_D = 12, RESULT_FIELDS = 13;

	public static final int METHOD_PEARSONS_CORRELATION = 0, METHOD_SIMPLE_SCORE = 1;
	public static final int JITTER_NAN = 0, JITTER_PHILOX = 1;
	public static final int PERM_REPLAY = 0, PERM_PHILOX = 1;
	public static final int KS_EXACT = 0, KS_FAST = 1;
	public static final int COMM_ID_BYTES = 128;

	private long ctx; // pctsea_ctx*
	private long matrix; // pctsea_matrix*

	public PctseaB200(int device) {
		ctx = create(device);
	}

	// ---- ctx -------------------------------------------------------------------------------------------
	private static native int abiVersion(); // pctsea_abi_version

	private static native long create(int device); // pctsea_create

	private static native void destroy(long ctx); // pctsea_destroy

	private static native String lastError(long ctx); // pctsea_last_error

	/** {setup, score, rank, observed, labels, null, reduce, total, gather} ms + {nPass, nUsed, launches, h2d, d2h} */
	private static native void getTimings(long ctx, float[] ms9, long[] counts5); // pctsea_get_timings

	private static native void setOption(long ctx, int option, long value); // pctsea_ctx_set_option

	// ---- matrix ----------------------------------------------------------------------------------------
	private static native long loadCsr(long ctx, long nCells, int nGenes, long[] rowPtr, int[] colIdx, float[] val); // pctsea_matrix_load_csr

	private static native long loadCsrDirect(long ctx, long nCells, int nGenes, java.nio.ByteBuffer rowPtr,
			java.nio.ByteBuffer colIdx, java.nio.ByteBuffer val); // pctsea_matrix_load_csr, direct buffers

	private static native void setLabels(long ctx, long matrix, int[] label, int nTypes); // pctsea_matrix_set_labels

	private static native void buildGeneIndex(long ctx, long matrix); // pctsea_matrix_build_gene_index

	private static native void freeMatrix(long ctx, long matrix); // pctsea_matrix_free

	// ---- stages ----------------------------------------------------------------------------------------
	private static native void score(long ctx, long matrix, int[] geneIdx, float[] abundance, int method,
			int minGenesCells, double minCorr, int jitterMode, long seed, double[] scoreOut, int[] nGenesUsedOut,
			int[] nNonZeroOut, byte[] keptOut); // pctsea_score

	private static native long rank(long ctx, double[] score, byte[] kept, double minScore, int[] orderOut); // pctsea_rank

	private static native void enrich(long ctx, double[] score, long nUsed, int[] order, int[] label, int nTypes,
			int nPerm, int permBegin, int permCount, int permMode, int ksMode, int feistelRounds, long seed,
			int meanPolicy, int[] replayPerm, double[][] result, float[] nullOut, float[] nullDOut); // pctsea_enrich

	private static native void reduceNull(long ctx, int nTypes, int nPerm, float[] nullEws, int meanPolicy,
			double[][] resultInOut); // pctsea_reduce_null

	// ---- whole path ------------------------------------------------------------------------------------
	private static native long analyze(long ctx, long matrix, int[] geneIdx, float[] abundance, int method,
			int minGenesCells, double minCorr, int jitterMode, long scoreSeed, double minScore, int[] label,
			int nTypes, int nPerm, int permBegin, int permCount, int permMode, int ksMode, int feistelRounds,
			long permSeed, int meanPolicy, int minPass, int[] replayPerm, double[][] result, float[] nullOut,
			float[] nullDOut, double[] scoreOut, int[] nGenesUsedOut, byte[] keptOut, int[] orderOut); // pctsea_analyze

	private static native void analyzeBatch(long ctx, long matrix, long[] listPtr, int[] geneIdx, float[] abundance,
			int method, int minGenesCells, double minCorr, int jitterMode, long scoreSeed, double minScore,
			int[] label, int nTypes, int nPerm, int ksMode, int feistelRounds, long permSeed, int meanPolicy,
			int minPass, double[][][] result, long[] nPassOut, int[] statusOut, float[] nullOut, float[] nullDOut); // pctsea_analyze_batch

	// ---- by-products of the last analyze ----------------------------------------------------------------
	private static native long lastTypeCounts(long ctx, int nTypes, int[] nCellsOfType, int[] nCellsOfTypePassing); // pctsea_last_type_counts

	private static native void lastObservedCurve(long ctx, int typeId, long nUsed, float[] aOut, float[] bOut); // pctsea_last_observed_curve

	// ---- multi-GPU --------------------------------------------------------------------------------------
	private static native void commUniqueId(byte[] idOut128); // pctsea_comm_unique_id

	private static native void commInitRank(long ctx, int nRanks, int rank, byte[] id128); // pctsea_comm_init_rank

	private static native void commDestroy(long ctx); // pctsea_comm_destroy

	// =====================================================================================================
	// public surface used by the patched PCTSEA.run() (java/patches/PCTSEA_run.patch)
	// =====================================================================================================
	public static int version() {
		return abiVersion();
	}

	/** Dataset loaded once and kept on the device (rows = cells in DB order, columns = dataset genes). */
	public void load(long[] rowPtr, int[] col, float[] val, int nGenes, int[] label, int nTypes, boolean geneIndex) {
		if (matrix != 0) {
			freeMatrix(ctx, matrix);
			matrix = 0;
		}
		matrix = loadCsr(ctx, rowPtr.length - 1, nGenes, rowPtr, col, val);
		setLabels(ctx, matrix, label, nTypes);
		if (geneIndex) {
			buildGeneIndex(ctx, matrix);
		}
	}

	public void loadDirect(long nCells, int nGenes, java.nio.ByteBuffer rowPtr, java.nio.ByteBuffer col,
			java.nio.ByteBuffer val, int[] label, int nTypes, boolean geneIndex) {
		if (matrix != 0) {
			freeMatrix(ctx, matrix);
			matrix = 0;
		}
		matrix = loadCsrDirect(ctx, nCells, nGenes, rowPtr, col, val);
		setLabels(ctx, matrix, label, nTypes);
		if (geneIndex) {
			buildGeneIndex(ctx, matrix);
		}
	}

	public void relabel(int[] label, int nTypes) {
		setLabels(ctx, matrix, label, nTypes);
	}

	/** Parameters of one round of PCTSEA.run() (one scoring schema). */
	public static final class Round {
		public int method = METHOD_PEARSONS_CORRELATION;
		public int minGenesCells = 4;
		public double minCorr = 0.1;
		public int jitterMode = JITTER_PHILOX;
		public long scoreSeed = 1L;
		public double minScore = 0.0;
		public int nPerm = 1000;
		public int permMode = PERM_PHILOX;
		public int ksMode = KS_FAST;
		public int feistelRounds = 0; // library default
		public long permSeed = 1L;
		public int meanPolicy = 0; // Maths.mean(TFloatList): fp32 running sum
		public int minPass = 1000; // PCTSEA.MIN_CELLS_PASSING_SCORE_TRESHOLD
	}

	/** Everything PCTSEA.run() consumes after the hot path. */
	public static final class Analysis {
		public double[][] result; // [nTypes][RESULT_FIELDS]
		public float[] nullEws, nullD; // [nTypes][nPerm] row-major
		public double[] score; // [nCells]
		public int[] nGenesUsed; // [nCells]
		public byte[] kept; // [nCells]
		public int[] order; // [nPass]; the last entry is the cell dropped by subList(0, size - 1)
		public long nPass;
	}

	public Analysis analyze(int nCells, int[] geneIdx, float[] abundance, int nTypes, Round r, boolean wantNull) {
		final Analysis a = new Analysis();
		a.result = new double[nTypes][RESULT_FIELDS];
		a.score = new double[nCells];
		a.nGenesUsed = new int[nCells];
		a.kept = new byte[nCells];
		a.order = new int[nCells];
		if (wantNull) {
			a.nullEws = new float[nTypes * r.nPerm];
			a.nullD = new float[nTypes * r.nPerm];
		}
		a.nPass = analyze(ctx, matrix, geneIdx, abundance, r.method, r.minGenesCells, r.minCorr, r.jitterMode,
				r.scoreSeed, r.minScore, null, nTypes, r.nPerm, 0, 0, r.permMode, r.ksMode, r.feistelRounds,
				r.permSeed, r.meanPolicy, r.minPass, null, a.result, a.nullEws, a.nullD, a.score, a.nGenesUsed, a.kept,
				a.order);
		return a;
	}

	/** -load_random (PCTSEA.java:1385-1387): a10-a12 from a null matrix read from the random-distributions file. */
	public void reduce(int nTypes, int nPerm, float[] nullEws, int meanPolicy, double[][] resultInOut) {
		reduceNull(ctx, nTypes, nPerm, nullEws, meanPolicy, resultInOut);
	}

	/** Inputs of calculateHyperGeometricStatistics (PCTSEA.java:2169-2264); returns N (cells kept). */
	public long typeCounts(int nTypes, int[] nCellsOfType, int[] nCellsOfTypePassing) {
		return lastTypeCounts(ctx, nTypes, nCellsOfType, nCellsOfTypePassing);
	}

	/** a / b series of writeScoreCalculationFile (EnrichmentWeigthedScoreParallel.java:355-410). */
	public void observedCurve(int typeId, long nUsed, float[] aOut, float[] bOut) {
		lastObservedCurve(ctx, typeId, nUsed, aOut, bOut);
	}

	/** Many lists against the resident matrix (AnalyzeView.java:747-766 submits one run per list). */
	public void analyzeBatch(long[] listPtr, int[] geneIdx, float[] abundance, int nTypes, Round r,
			double[][][] result, long[] nPass, int[] status) {
		analyzeBatch(ctx, matrix, listPtr, geneIdx, abundance, r.method, r.minGenesCells, r.minCorr, r.jitterMode,
				r.scoreSeed, r.minScore, null, nTypes, r.nPerm, r.ksMode, r.feistelRounds, r.permSeed, r.meanPolicy,
				r.minPass, result, nPass, status, null, null);
	}

	/** Rank 0 creates the id, the host hands it to the other ranks (threads of this JVM, or processes). */
	public static byte[] newCommunicatorId() {
		final byte[] id = new byte[COMM_ID_BYTES];
		commUniqueId(id);
		return id;
	}

	/** Collective: every rank's PctseaB200 calls it with the same id. analyze() then shards the permutations. */
	public void joinCommunicator(int nRanks, int rank, byte[] id) {
		commInitRank(ctx, nRanks, rank, id);
	}

	public void leaveCommunicator() {
		commDestroy(ctx);
	}

	public void option(int option, long value) {
		setOption(ctx, option, value);
	}

	public String timingsLine() {
		final float[] ms = new float[9];
		final long[] n = new long[5];
		getTimings(ctx, ms, n);
		return String.format("score %.3f ms, rank %.3f ms, observed %.3f ms, null %.3f ms, gather %.3f ms, reduce %.3f ms, "
				+ "total %.3f ms (%d cells ranked, %d kernel launches)", ms[1], ms[2], ms[3], ms[5], ms[8], ms[6], ms[7],
				n[1], n[2]);
	}

	@Override
	public void close() {
		if (matrix != 0) {
			freeMatrix(ctx, matrix);
			matrix = 0;
		}
		if (ctx != 0) {
			destroy(ctx);
			ctx = 0;
		}
	}
}
