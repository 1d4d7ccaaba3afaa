# A/B of the two scoring paths on config B: per-step device totals (median, mean, max) and stage times
for rep in 1 2 3; do for gi in 0 1; do
  echo -n "gene_index=$gi rep=$rep "; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gene-index $gi 2>&1 | tail -1 | python -c "import sys,json,statistics as st; d=json.loads(sys.stdin.read()); t=d['per_step_total_ms']; s=d['stage_ms']; print('value %.4g' % d['value'], 'e2e_ms %.3f' % d['ms_per_step'], 'dev median %.3f mean %.3f max %.3f' % (st.median(t), sum(t)/len(t), max(t)), 'score %.3f rank %.3f labels %.3f null %.3f' % (s['score_ms'], s['rank_ms'], s['labels_ms'], s['null_ms']))"
done; done
