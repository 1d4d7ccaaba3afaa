import java.io.File;
import java.io.FileWriter;
import java.io.IOException;
import java.io.PrintWriter;
import java.nio.file.Files;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.Collections;
import java.util.List;
import java.util.Random;

import org.apache.commons.math3.stat.correlation.PearsonsCorrelation;
import org.apache.commons.math3.stat.inference.KolmogorovSmirnovTest;

import edu.scripps.yates.pctsea.model.CellTypeClassification;
import edu.scripps.yates.pctsea.model.ScoringMethod;
import edu.scripps.yates.pctsea.model.SingleCell;
import edu.scripps.yates.pctsea.utils.PCTSEAUtils;
import edu.scripps.yates.pctsea.utils.parallel.EnrichmentWeigthedScoreParallel;
import edu.scripps.yates.utilities.pi.ParIterator;
import edu.scripps.yates.utilities.pi.ParIterator.Schedule;
import edu.scripps.yates.utilities.pi.ParIteratorFactory;

/**
 * Pins the CPU oracle (oracle/pctsea_oracle.c) to the REAL reference: runs the reference's own classes and the JDK /
 * commons-math calls its hot path makes on small fixed inputs, and writes inputs + outputs (floating point as IEEE bit
 * patterns) to golden_ref.json. tests/test_oracle_cpu.py::test_reference_pinned_golden_vectors replays the inputs
 * through the oracle and compares bit for bit; while the file is absent the suite reports "parity unpinned".
 *
 * It has NOT been compiled or run by the authors of this repository (no JDK in their image; pctsea-core depends on
 * private -SNAPSHOT artifacts). Build and run next to a built pctsea-core, see README.md in this folder.
 *
 * What is exercised (reference file:line):
 *   jdk_random / shuffle   Collections.shuffle(list, rnd) as in PCTSEA.java:1405 (there with an unseeded Random)
 *   pearson                PearsonsCorrelation.correlation as in model/SingleCell.java:449-450
 *   sort_search            Arrays.sort(float[]) + Arrays.binarySearch as in PCTSEA.java:1523-1537 (duplicate keys)
 *   ks_statistic           KolmogorovSmirnovTest.kolmogorovSmirnovStatistic as in ...Parallel.java:211-212
 *   enrichment             EnrichmentWeigthedScoreParallel.run, observed (:72-353) and permuted (:315) data
 *   normalized             CellTypeClassification.getNormalizedEnrichmentScore (model/CellTypeClassification.java:407-441)
 */
public class DumpGolden {
	static String f64(double v) {
		return "\"" + Long.toHexString(Double.doubleToRawLongBits(v)) + "\"";
	}

	static String f32(float v) {
		return "\"" + Integer.toHexString(Float.floatToRawIntBits(v)) + "\"";
	}

	static String arr(List<String> items) {
		return "[" + String.join(",", items) + "]";
	}

	public static void main(String[] args) throws Exception {
		final File out = new File(args.length > 0 ? args[0] : "golden_ref.json");
		final List<String> sections = new ArrayList<String>();

		// ---- java.util.Random + Collections.shuffle
		{
			final List<String> parts = new ArrayList<String>();
			for (final long seed : new long[] { 42L, 0L }) {
				final Random r = new Random(seed);
				final List<String> ints = new ArrayList<String>(), bounded = new ArrayList<String>(),
						dbl = new ArrayList<String>();
				for (int i = 0; i < 8; i++) {
					ints.add(Integer.toString(r.nextInt()));
				}
				for (final int bound : new int[] { 10, 16, 1000, 219530, 1 << 30, (1 << 30) + 1 }) {
					bounded.add(Integer.toString(r.nextInt(bound)));
				}
				for (int i = 0; i < 4; i++) {
					dbl.add(f64(r.nextDouble()));
				}
				parts.add("{\"seed\":" + seed + ",\"nextInt\":" + arr(ints)
						+ ",\"nextIntBounds\":[10,16,1000,219530,1073741824,1073741825],\"nextIntBounded\":" + arr(bounded)
						+ ",\"nextDouble\":" + arr(dbl) + "}");
			}
			sections.add("\"jdk_random\":" + arr(parts));
			final List<String> sh = new ArrayList<String>();
			for (final int n : new int[] { 10, 257 }) {
				final List<Integer> list = new ArrayList<Integer>();
				for (int i = 0; i < n; i++) {
					list.add(i);
				}
				Collections.shuffle(list, new Random(42));
				final List<String> s = new ArrayList<String>();
				for (final Integer v : list) {
					s.add(v.toString());
				}
				sh.add("{\"n\":" + n + ",\"seed\":42,\"result\":" + arr(s) + "}");
			}
			sections.add("\"shuffle\":" + arr(sh));
		}

		// ---- PearsonsCorrelation.correlation(interactors, genes)
		{
			final Random r = new Random(3);
			final List<String> cases = new ArrayList<String>();
			for (final int n : new int[] { 2, 3, 5, 17, 120 }) {
				final double[] x = new double[n], y = new double[n];
				for (int i = 0; i < n; i++) {
					x[i] = (float) (1 + r.nextInt(40)); // abundances: small integers as floats
					y[i] = (float) (1 + r.nextInt(12)); // UMI counts
				}
				final double rho = new PearsonsCorrelation().correlation(x, y);
				final List<String> xs = new ArrayList<String>(), ys = new ArrayList<String>();
				for (int i = 0; i < n; i++) {
					xs.add(f64(x[i]));
					ys.add(f64(y[i]));
				}
				cases.add("{\"x\":" + arr(xs) + ",\"y\":" + arr(ys) + ",\"r\":" + f64(rho) + "}");
			}
			sections.add("\"pearson\":" + arr(cases));
		}

		// ---- Arrays.sort(float[]) and Arrays.binarySearch with duplicates, -0.0f / 0.0f and NaN
		{
			final float[] a = { 0.5f, -0.0f, 0.0f, 2.0f, 0.5f, 0.5f, Float.NaN, 1.25f, 0.5f, -3.0f, 2.0f, 0.5f };
			final float[] s = a.clone();
			Arrays.sort(s);
			final List<String> in = new ArrayList<String>(), sorted = new ArrayList<String>(), res = new ArrayList<String>();
			for (final float v : a) {
				in.add(f32(v));
			}
			for (final float v : s) {
				sorted.add(f32(v));
			}
			final float[] keys = { 0.5f, 2.0f, 0.0f, -0.0f, 1.0f, 9.0f, -9.0f, Float.NaN };
			for (final float k : keys) {
				res.add("{\"key\":" + f32(k) + ",\"index\":" + Arrays.binarySearch(s, k) + "}");
			}
			sections.add("\"sort_search\":{\"input\":" + arr(in) + ",\"sorted\":" + arr(sorted) + ",\"searches\":"
					+ arr(res) + "}");
		}

		// ---- ranked list shared by the KS sections: 400 cells, 5 types + unknown, scores descending, with ties
		final int n = 400, T = 5;
		final Random rg = new Random(11);
		final double[] score = new double[n];
		final int[] type = new int[n];
		for (int i = 0; i < n; i++) {
			score[i] = 0.2 + 0.8 * Math.round(rg.nextDouble() * 64) / 64.0; // ties on purpose
			final int t = rg.nextInt(T + 1);
			type[i] = t == T ? -1 : (rg.nextInt(4) == 0 ? 0 : t); // type 0 over-represented
		}
		final List<SingleCell> cells = new ArrayList<SingleCell>();
		for (int i = 0; i < n; i++) {
			final SingleCell c = new SingleCell(i + 1, "c" + i, score[i]);
			c.setCellType(type[i] < 0 ? null : "golden type " + type[i], false, null);
			cells.add(c);
		}
		PCTSEAUtils.sortByScoreDescending(cells); // the order the KS stage sees (utils/PCTSEAUtils.java:44-53)
		final List<String> rs = new ArrayList<String>(), rt = new ArrayList<String>();
		for (final SingleCell c : cells) {
			rs.add(f64(c.getScoreForRanking()));
			rt.add(c.getCellType() == null ? "-1" : c.getCellType().substring("golden type ".length()));
		}

		// ---- KolmogorovSmirnovTest.kolmogorovSmirnovStatistic: type 0 against the others
		{
			final List<Double> a = new ArrayList<Double>(), b = new ArrayList<Double>();
			for (final SingleCell c : cells) {
				("golden type 0".equals(c.getCellType()) ? a : b).add(c.getScoreForRanking());
			}
			final double[] x = a.stream().mapToDouble(Double::doubleValue).toArray();
			final double[] y = b.stream().mapToDouble(Double::doubleValue).toArray();
			sections.add("\"ks_statistic\":{\"type\":0,\"d\":" + f64(new KolmogorovSmirnovTest().kolmogorovSmirnovStatistic(x, y)) + "}");
		}

		// ---- EnrichmentWeigthedScoreParallel.run: observed, then 6 permutations (Collections.shuffle of the types)
		{
			final List<CellTypeClassification> cts = new ArrayList<CellTypeClassification>();
			for (int t = 0; t < T; t++) {
				cts.add(new CellTypeClassification("golden type " + t, Double.NaN));
			}
			final File tmp = Files.createTempDirectory("dumpgolden").toFile();
			runKs(cts, cells, false, tmp);
			final List<String> obs = new ArrayList<String>();
			for (final CellTypeClassification ct : cts) {
				obs.add("{\"type\":" + ct.getName().substring("golden type ".length()) + ",\"ews\":" + f32(ct.getEnrichmentScore())
						+ ",\"supX\":" + ct.getSupremumX() + ",\"norm_supX\":" + f64(ct.getNormalizedSupremumX())
						+ ",\"sizeA\":" + ct.getSizeA() + ",\"sizeB\":" + ct.getSizeB() + ",\"dstat\":" + f32(ct.getKSTestDStatistic())
						+ ",\"sec_ews\":" + (ct.getSecondaryEnrichmentScore() == null ? "null" : f32(ct.getSecondaryEnrichmentScore()))
						+ ",\"sec_supX\":" + (ct.getSecondarySupremumX() == null ? "null" : ct.getSecondarySupremumX().toString())
						+ ",\"ks_d\":" + f64(ct.getKSTestPvalue()) + "}");
			}
			// permutations as PCTSEA.java:1398-1429: shuffle the list of type names, reassign, rerun with permutatedData
			final List<String> original = new ArrayList<String>();
			cells.forEach(c -> original.add(c.getCellType()));
			final Random rp = new Random(7);
			final int P = 6;
			final List<String> perms = new ArrayList<String>();
			for (int p = 0; p < P; p++) {
				final List<String> shuffled = new ArrayList<String>(original);
				Collections.shuffle(shuffled, rp);
				final List<String> lab = new ArrayList<String>();
				for (int i = 0; i < cells.size(); i++) {
					cells.get(i).setCellType(shuffled.get(i), false, null);
					lab.add(shuffled.get(i) == null ? "-1" : shuffled.get(i).substring("golden type ".length()));
				}
				runKs(cts, cells, true, tmp);
				perms.add(arr(lab));
			}
			for (int i = 0; i < cells.size(); i++) {
				cells.get(i).setCellType(original.get(i), false, null);
			}
			final List<String> nul = new ArrayList<String>(), norm = new ArrayList<String>();
			for (final CellTypeClassification ct : cts) {
				final List<String> e = new ArrayList<String>(), d = new ArrayList<String>();
				for (int p = 0; p < ct.getRandomEnrichmentScores().size(); p++) {
					e.add(f32(ct.getRandomEnrichmentScores().get(p)));
					d.add(f32(ct.getRandomKSTestDStatistics().get(p)));
				}
				nul.add("{\"ews\":" + arr(e) + ",\"dstat\":" + arr(d) + "}");
				norm.add(f32(ct.getNormalizedEnrichmentScore())); // model/CellTypeClassification.java:407-441
			}
			sections.add("\"enrichment\":{\"n_types\":" + T + ",\"score\":" + arr(rs) + ",\"type\":" + arr(rt)
					+ ",\"observed\":" + arr(obs) + ",\"perm_seed\":7,\"perm_labels\":" + arr(perms) + ",\"null\":" + arr(nul)
					+ ",\"normalized\":" + arr(norm) + "}");
		}

		try (PrintWriter w = new PrintWriter(new FileWriter(out))) {
			w.println("{\"generator\":\"tests/golden/java/DumpGolden.java\",\"java\":\"" + System.getProperty("java.version")
					+ "\"," + String.join(",\n", sections) + "}");
		}
		System.out.println("wrote " + out.getAbsolutePath());
	}

	/** One thread, the reference's own runner (PCTSEA.java:2490-2519 starts 2 x cores of them). */
	static void runKs(List<CellTypeClassification> cts, List<SingleCell> ranked, boolean permuted, File tmp)
			throws InterruptedException {
		final ParIterator<CellTypeClassification> it = ParIteratorFactory.createParIterator(cts, 1, Schedule.GUIDED);
		final EnrichmentWeigthedScoreParallel runner = new EnrichmentWeigthedScoreParallel(it, 1, ranked, permuted, false,
				ScoringMethod.PEARSONS_CORRELATION.getScoreName(), tmp, "golden", true, ScoringMethod.PEARSONS_CORRELATION,
				null);
		runner.start();
		runner.join();
	}
}
