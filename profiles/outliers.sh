for iv in 5 25 100 none; do for rep in 1 2 3; do
  if [ $iv = none ]; then export PCTSEA_BENCH_NO_SAMPLER=1; else unset PCTSEA_BENCH_NO_SAMPLER; export PCTSEA_BENCH_SAMPLE_MS=$iv; fi
  echo -n "interval=$iv rep=$rep "; timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json,statistics as st; d=json.loads(sys.stdin.read()); t=d['per_step_total_ms']; print('e2e_ms', round(d['ms_per_step'],3), 'dev_mean', round(sum(t)/len(t),3), 'median', round(st.median(t),3), 'max', max(t), 'n>4.2:', sum(x>4.2 for x in t), 'samples', (d.get('clocks') or {}).get('samples'))"
done; done
