package edu.scripps.yates.pctsea.gpu;

import java.io.BufferedReader;
import java.io.DataInputStream;
import java.io.File;
import java.io.FileInputStream;
import java.io.FileReader;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.channels.FileChannel;
import java.util.HashMap;
import java.util.Iterator;
import java.util.List;
import java.util.Map;

import edu.scripps.yates.pctsea.GeneExpressionsRetriever;
import edu.scripps.yates.pctsea.model.CellTypeClassification;
import edu.scripps.yates.pctsea.model.CellTypes;
import edu.scripps.yates.pctsea.model.ScoringMethod;
import edu.scripps.yates.pctsea.model.ScoringSchema;
import edu.scripps.yates.pctsea.model.SingleCell;
import gnu.trove.list.TShortList;

/**
 * The single-cell dataset as libpctsea_b200.so wants it: a CSR matrix written once by MongoToCsr (cells in
 * findByDatasetTag order = the cellID order of PCTSEA.getSingleCellListFromDB, PCTSEA.java:1316,1347-1355; genes =
 * upper-cased names), loaded once onto the GPU and kept there for every analysis against that dataset — the role the
 * MongoDB "expression" collection plays for GeneExpressionsRetriever.getSingleCellExpressionsFromDB
 * (GeneExpressionsRetriever.java:120-274), without the per-list streaming.
 *
 * This class is the host-side glue of java/patches/PCTSEA_run.patch: it maps the input list to matrix columns, calls
 * PctseaB200.analyze and writes the results back into the objects the rest of PCTSEA.run() consumes. It is NOT compiled
 * in the image this repository is developed in (no JDK, and pctsea-core's own dependencies are private snapshots);
 * the executable contract is the C ABI test-suite.
 *
 * File layout of "<base>.csr" written by MongoToCsr and read here (little endian):
 *   magic "PCTSEACSR1\0\0" (12 bytes) | nCells int64 | nGenes int32 | nTypes int32 | nnz int64 |
 *   rowPtr int64[nCells + 1] | col int32[nnz] | val float32[nnz] | label int32[nCells]
 * and "<base>.genes.txt" / "<base>.types.txt" / "<base>.cells.txt": one name per line, line i = column / label id / row i.
 * pctsea_parent_b200/ingest.py reads and writes the same files (CSRDataset.load_binary / save_binary).
 */
public final class GpuDataset implements AutoCloseable {
	private final PctseaB200 gpu;
	private final Map<String, Integer> columnByGene = new HashMap<String, Integer>();
	private final String[] typeNames;
	private final int nCells, nGenes, nTypes;
	private PctseaB200.Analysis last;
	private PctseaB200.Round lastRound;

	public GpuDataset(int device, File base, boolean geneIndex) throws IOException {
		final String[] genes = readLines(new File(base.getPath() + ".genes.txt"));
		typeNames = readLines(new File(base.getPath() + ".types.txt"));
		for (int g = 0; g < genes.length; g++) {
			columnByGene.put(genes[g].toUpperCase(), g);
		}
		gpu = new PctseaB200(device);
		try (FileInputStream in = new FileInputStream(base.getPath() + ".csr"); FileChannel ch = in.getChannel()) {
			final ByteBuffer head = ByteBuffer.allocate(12 + 8 + 4 + 4 + 8).order(ByteOrder.LITTLE_ENDIAN);
			ch.read(head);
			head.flip();
			final byte[] magic = new byte[12];
			head.get(magic);
			if (!new String(magic, 0, 10, "US-ASCII").equals("PCTSEACSR1")) {
				throw new IOException(base + ".csr is not a PCTSEA CSR dump");
			}
			nCells = (int) head.getLong();
			nGenes = head.getInt();
			nTypes = head.getInt();
			final long nnz = head.getLong();
			// direct buffers: no 2^31 array limit (an atlas-scale matrix has 3e9 non-zeros), no copy inside the JVM
			final ByteBuffer rowPtr = readDirect(ch, (nCells + 1L) * 8);
			final ByteBuffer col = readDirect(ch, nnz * 4);
			final ByteBuffer val = readDirect(ch, nnz * 4);
			final ByteBuffer lab = readDirect(ch, nCells * 4L);
			final int[] label = new int[nCells];
			lab.asIntBuffer().get(label);
			gpu.loadDirect(nCells, nGenes, rowPtr, col, val, label, nTypes, geneIndex);
		}
		if (typeNames.length != nTypes || genes.length != nGenes) {
			throw new IOException("name files do not match the CSR header");
		}
	}

	private static ByteBuffer readDirect(FileChannel ch, long bytes) throws IOException {
		if (bytes > Integer.MAX_VALUE) {
			// one mapping per 1 GiB would be needed for a contiguous native block above 2 GiB: MongoToCsr splits such
			// arrays over "<base>.csr.partN" files; out of scope of this sketch of the host glue
			throw new IOException("array above 2 GiB: load it part by part (see MongoToCsr)");
		}
		final ByteBuffer b = ByteBuffer.allocateDirect((int) bytes).order(ByteOrder.nativeOrder());
		while (b.hasRemaining() && ch.read(b) >= 0) {
		}
		b.flip();
		return b;
	}

	private static String[] readLines(File f) throws IOException {
		final java.util.ArrayList<String> out = new java.util.ArrayList<String>();
		try (BufferedReader r = new BufferedReader(new FileReader(f))) {
			String line;
			while ((line = r.readLine()) != null) {
				out.add(line);
			}
		}
		return out.toArray(new String[0]);
	}

	/** Stages 1-3 for one scoring schema (the body of the per-schema loop, PCTSEA.java:348-434). */
	public PctseaB200.Analysis analyze(GeneExpressionsRetriever list, ScoringSchema schema, Double minCorr, int nPerm,
			int minPass, long seed, boolean wantNull) {
		final TShortList geneIDs = list.getGeneIDs();
		final int[] geneIdx = new int[geneIDs.size()];
		final float[] abundance = new float[geneIDs.size()];
		for (int l = 0; l < geneIdx.length; l++) {
			final short id = geneIDs.get(l);
			final Integer column = columnByGene.get(list.getGeneName(id).toUpperCase());
			geneIdx[l] = column == null ? -1 : column; // -1: gene absent from the dataset
			abundance[l] = list.getGeneExpressionInOurExperiment(id);
		}
		final PctseaB200.Round r = new PctseaB200.Round();
		r.method = schema.getScoringMethod() == ScoringMethod.PEARSONS_CORRELATION
				? PctseaB200.METHOD_PEARSONS_CORRELATION
				: PctseaB200.METHOD_SIMPLE_SCORE;
		r.minGenesCells = schema.getMinGenesCells();
		r.minCorr = minCorr; // a null Double NPEs in the reference as well (PCTSEA.java:379-381)
		r.minScore = schema.getScoringThreshold().getThresholdValue();
		r.nPerm = nPerm;
		r.minPass = minPass;
		r.scoreSeed = seed;
		r.permSeed = seed;
		last = gpu.analyze(nCells, geneIdx, abundance, nTypes, r, wantNull);
		lastRound = r;
		return last;
	}

	public PctseaB200.Analysis lastAnalysis() {
		return last;
	}

	/** SingleCell.setScore for every cell (model/SingleCell.java:315); cell ids are 1-based DB order. */
	public void writeScores(PctseaB200.Analysis a, List<SingleCell> cells) {
		for (final SingleCell c : cells) {
			c.setScore(a.score[c.getId() - 1]);
		}
	}

	/** The min_genes_cells filter (PCTSEA.java:544-572) = the kept flags of the same GPU call. */
	public void retainKept(PctseaB200.Analysis a, List<SingleCell> cellsOfThisRound) {
		final Iterator<SingleCell> it = cellsOfThisRound.iterator();
		while (it.hasNext()) {
			if (a.kept[it.next().getId() - 1] == 0) {
				it.remove();
			}
		}
	}

	/**
	 * What calculateEnrichmentScore + calculateSignificanceByCellTypesPermutations leave in the
	 * CellTypeClassification objects (PCTSEA.java:427-434, 1380-1549).
	 */
	public void fillCellTypeClassifications(List<CellTypeClassification> cellTypes, File randomScoresFileOrNull)
			throws IOException {
		final int P = lastRound.nPerm;
		float[] nullEws = last.nullEws, nullD = last.nullD;
		if (nullEws == null) {
			// -load_random (PCTSEA.java:1385-1387, 1834-1869): null from the file, a10-a12 by pctsea_reduce_null
			nullEws = new float[nTypes * P];
			nullD = new float[nTypes * P];
			java.util.Arrays.fill(nullEws, Float.NaN);
			readRandomDistributions(randomScoresFileOrNull, nullEws, nullD, P);
			gpu.reduce(nTypes, P, nullEws, lastRound.meanPolicy, last.result);
		}
		for (final CellTypeClassification ct : cellTypes) {
			final int t = labelOf(ct.getName());
			if (t < 0) {
				continue;
			}
			final double[] r = last.result[t];
			ct.setEnrichment((float) r[PctseaB200.EWS], (float) r[PctseaB200.DSTAT], (int) r[PctseaB200.SUP_X],
					r[PctseaB200.NORM_SUP_X]);
			ct.setSizeA((int) r[PctseaB200.SIZE_A]);
			ct.setSizeB((int) r[PctseaB200.SIZE_B]);
			for (int p = 0; p < P; p++) {
				ct.addRandomEnrichment(nullEws[t * P + p], nullD[t * P + p]);
			}
			ct.setNormalizedEnrichmentScore((float) r[PctseaB200.NORM_EWS]);
			ct.setEnrichmentSignificance(r[PctseaB200.P_EMP]);
			ct.setEnrichmentFDR(r[PctseaB200.FDR]);
			if (r[PctseaB200.SEC_SUP_X] != 0) { // EnrichmentWeigthedScoreParallel.java:308-310
				ct.setSecondaryEnrichment((float) r[PctseaB200.SEC_EWS], (int) r[PctseaB200.SEC_SUP_X]);
			}
			ct.setKSTestPvalue(r[PctseaB200.KS_D]); // ...Parallel.java:228
			ct.setEnrichmentUnweightedScore((float) r[PctseaB200.KS_D]); // PCTSEA.java:2616
		}
	}

	private int labelOf(String typeName) {
		for (int t = 0; t < typeNames.length; t++) {
			if (typeNames[t].equals(typeName)) {
				return t;
			}
		}
		return -1;
	}

	private void readRandomDistributions(File f, float[] nullEws, float[] nullD, int P) throws IOException {
		try (BufferedReader reader = new BufferedReader(new FileReader(f))) {
			String line = reader.readLine(); // header
			while ((line = reader.readLine()) != null) {
				final String[] split = line.split("\t");
				int t = labelOf(split[0]);
				if (t < 0) {
					t = labelOf(SingleCell.parseCellTypeTypos(split[0]));
				}
				if (t < 0) {
					continue;
				}
				for (int i = 5, p = 0; i + 1 < split.length && p < P; i += 2, p++) {
					nullEws[t * P + p] = Float.valueOf(split[i]);
					nullD[t * P + p] = Float.valueOf(split[i + 1]);
				}
			}
		}
	}

	/** Labels for another branch of the cell-type tree (CellTypeBranch != ORIGINAL): re-upload, matrix stays. */
	public void relabel(List<SingleCell> cellsInDbOrder) {
		final int[] label = new int[nCells];
		java.util.Arrays.fill(label, -1);
		for (final SingleCell c : cellsInDbOrder) {
			label[c.getId() - 1] = c.getCellTypeID() < 0 ? -1 : labelOf(c.getCellType());
		}
		gpu.relabel(label, nTypes);
	}

	public int getNumCells() {
		return nCells;
	}

	@Override
	public void close() {
		gpu.close();
	}
}
