package edu.scripps.yates.pctsea.gpu;

import java.io.BufferedOutputStream;
import java.io.BufferedWriter;
import java.io.DataOutputStream;
import java.io.File;
import java.io.FileOutputStream;
import java.io.FileWriter;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

import org.bson.Document;
import org.springframework.data.mongodb.core.MongoTemplate;
import org.springframework.data.mongodb.core.query.Criteria;
import org.springframework.data.mongodb.core.query.Query;

import edu.scripps.yates.pctsea.db.SingleCellMongoRepository;
import edu.scripps.yates.pctsea.model.SingleCell;

/**
 * SURVEY.md §8 f1: dump one dataset of the reference's MongoDB into the CSR files GpuDataset (and
 * pctsea_parent_b200/ingest.py: CSRDataset.load_binary) read — once per dataset, instead of streaming the "expression"
 * collection for every analysis (MongoBaseService.getExpressionByGenes, db/MongoBaseService.java:341-355).
 *
 * Conventions, each taken from the code path it replaces:
 *  - rows = cells in the order singleCellMongoRepo.findByDatasetTag returns them (PCTSEA.java:1316); the reference gives
 *    them cellID 1, 2, ... in that order (PCTSEA.java:1347-1355), and ties of the ranking keep that order;
 *  - a cell's label = the id of its (typo-parsed) cell type, -1 when the document has no type
 *    (model/SingleCell.java:212-229); cells without a type never receive expressions
 *    (GeneExpressionsRetriever.java:149-153) but stay in the list;
 *  - columns = distinct upper-cased gene names (GeneExpressionsRetriever.java:177) in order of first appearance;
 *  - an expression document with value 0 is skipped (Float.compare(0f, expression) == 0, :170-172); a NaN
 *    expression (null field) is kept — the scoring kernels treat it like the reference does (x > 0 is false,
 *    x != 0 is true);
 *  - a (cell, gene) pair seen twice keeps the LAST document (model/SingleCell.java:80).
 *
 * Usage (inside the Spring context of pctsea-cl, which provides the MongoTemplate and the repository):
 *   new MongoToCsr(mongoTemplate, singleCellMongoRepo).dump("HCL", new File("/data/hcl"));
 * Not compiled in the image this repository is developed in (no JDK, no MongoDB); the Python twin of the writer and
 * reader is tested (tests/test_oracle_cpu.py::test_hcl_ingestion_to_csr).
 */
public final class MongoToCsr {
	private final MongoTemplate mongoTemplate;
	private final SingleCellMongoRepository singleCellMongoRepo;

	public MongoToCsr(MongoTemplate mongoTemplate, SingleCellMongoRepository singleCellMongoRepo) {
		this.mongoTemplate = mongoTemplate;
		this.singleCellMongoRepo = singleCellMongoRepo;
	}

	public void dump(String datasetTag, File base) throws IOException {
		// ---- cells, in DB order
		final List<edu.scripps.yates.pctsea.db.SingleCell> cells = singleCellMongoRepo.findByDatasetTag(datasetTag);
		final int nCells = cells.size();
		final Map<String, Integer> rowByCellName = new HashMap<String, Integer>(2 * nCells);
		final Map<String, Integer> labelByType = new HashMap<String, Integer>();
		final List<String> typeNames = new ArrayList<String>();
		final int[] label = new int[nCells];
		for (int i = 0; i < nCells; i++) {
			final edu.scripps.yates.pctsea.db.SingleCell c = cells.get(i);
			rowByCellName.put(c.getName(), i);
			if (c.getType() == null) {
				label[i] = -1;
				continue;
			}
			final String type = SingleCell.parseCellTypeTypos(c.getType());
			Integer id = labelByType.get(type);
			if (id == null) {
				id = typeNames.size();
				labelByType.put(type, id);
				typeNames.add(type);
			}
			label[i] = id;
		}

		// ---- pass 1 over the expression collection: row lengths and the gene dictionary
		final Map<String, Integer> columnByGene = new HashMap<String, Integer>();
		final List<String> geneNames = new ArrayList<String>();
		final long[] rowPtr = new long[nCells + 1];
		final Query query = Query.query(Criteria.where("projectTag").is(datasetTag));
		mongoTemplate.executeQuery(query, "expression", (Document doc) -> {
			final Object type = doc.get("cellType");
			final Double e = (Double) doc.get("expression");
			if (type == null || (e != null && Float.compare(0f, e.floatValue()) == 0)) {
				return;
			}
			final Integer row = rowByCellName.get((String) doc.get("cellName"));
			if (row == null) {
				return;
			}
			final String gene = ((String) doc.get("gene")).toUpperCase();
			if (!columnByGene.containsKey(gene)) {
				columnByGene.put(gene, geneNames.size());
				geneNames.add(gene);
			}
			rowPtr[row + 1]++;
		});
		for (int i = 0; i < nCells; i++) {
			rowPtr[i + 1] += rowPtr[i];
		}
		final long nnzUpper = rowPtr[nCells];
		if (nnzUpper > Integer.MAX_VALUE - 8) {
			throw new IOException("more than 2^31 non-zeros: dump the dataset in cell ranges and concatenate the parts");
		}

		// ---- pass 2: fill; a repeated (cell, gene) overwrites (last document wins)
		final int[] col = new int[(int) nnzUpper];
		final float[] val = new float[(int) nnzUpper];
		final long[] cursor = Arrays.copyOf(rowPtr, nCells);
		final List<Map<Integer, Integer>> slotOf = new ArrayList<Map<Integer, Integer>>(nCells);
		for (int i = 0; i < nCells; i++) {
			slotOf.add(null);
		}
		mongoTemplate.executeQuery(query, "expression", (Document doc) -> {
			final Object type = doc.get("cellType");
			final Double e = (Double) doc.get("expression");
			final float x = e == null ? Float.NaN : e.floatValue();
			if (type == null || Float.compare(0f, x) == 0) {
				return;
			}
			final Integer row = rowByCellName.get((String) doc.get("cellName"));
			if (row == null) {
				return;
			}
			final int g = columnByGene.get(((String) doc.get("gene")).toUpperCase());
			Map<Integer, Integer> slots = slotOf.get(row);
			if (slots == null) {
				slots = new HashMap<Integer, Integer>();
				slotOf.set(row, slots);
			}
			final Integer seen = slots.get(g);
			if (seen != null) {
				val[seen] = x;
				return;
			}
			final int k = (int) cursor[row]++;
			slots.put(g, k);
			col[k] = g;
			val[k] = x;
		});

		// ---- compact (repeated pairs left holes at the row ends), sort every row by column, write
		final long[] outPtr = new long[nCells + 1];
		int w = 0;
		for (int i = 0; i < nCells; i++) {
			final int b = (int) rowPtr[i], e = (int) cursor[i];
			final Integer[] order = new Integer[e - b];
			for (int k = 0; k < order.length; k++) {
				order[k] = b + k;
			}
			Arrays.sort(order, (x, y) -> Integer.compare(col[x], col[y]));
			final int[] c2 = new int[order.length];
			final float[] v2 = new float[order.length];
			for (int k = 0; k < order.length; k++) {
				c2[k] = col[order[k]];
				v2[k] = val[order[k]];
			}
			System.arraycopy(c2, 0, col, w, c2.length);
			System.arraycopy(v2, 0, val, w, v2.length);
			w += c2.length;
			outPtr[i + 1] = w;
			slotOf.set(i, null);
		}
		write(base, nCells, geneNames.size(), typeNames.size(), outPtr, col, val, w, label);
		writeLines(new File(base.getPath() + ".genes.txt"), geneNames);
		writeLines(new File(base.getPath() + ".types.txt"), typeNames);
		final List<String> cellNames = new ArrayList<String>(nCells);
		for (final edu.scripps.yates.pctsea.db.SingleCell c : cells) {
			cellNames.add(c.getName());
		}
		writeLines(new File(base.getPath() + ".cells.txt"), cellNames);
	}

	private static void write(File base, int nCells, int nGenes, int nTypes, long[] rowPtr, int[] col, float[] val,
			int nnz, int[] label) throws IOException {
		try (DataOutputStream out = new DataOutputStream(
				new BufferedOutputStream(new FileOutputStream(base.getPath() + ".csr"), 1 << 20))) {
			out.write(Arrays.copyOf("PCTSEACSR1".getBytes("US-ASCII"), 12));
			final ByteBuffer b = ByteBuffer.allocate(1 << 20).order(ByteOrder.LITTLE_ENDIAN);
			b.putLong(nCells).putInt(nGenes).putInt(nTypes).putLong(nnz);
			for (int i = 0; i <= nCells; i++) {
				flushIfFull(out, b);
				b.putLong(rowPtr[i]);
			}
			for (int k = 0; k < nnz; k++) {
				flushIfFull(out, b);
				b.putInt(col[k]);
			}
			for (int k = 0; k < nnz; k++) {
				flushIfFull(out, b);
				b.putFloat(val[k]);
			}
			for (int i = 0; i < nCells; i++) {
				flushIfFull(out, b);
				b.putInt(label[i]);
			}
			out.write(b.array(), 0, b.position());
		}
	}

	private static void flushIfFull(DataOutputStream out, ByteBuffer b) throws IOException {
		if (b.remaining() < 16) {
			out.write(b.array(), 0, b.position());
			b.clear();
		}
	}

	private static void writeLines(File f, List<String> lines) throws IOException {
		try (BufferedWriter wr = new BufferedWriter(new FileWriter(f))) {
			for (final String s : lines) {
				wr.write(s);
				wr.write('\n');
			}
		}
	}
}
