# Does the host scheduling class of the launching process change the step-time outliers? (box load is external)
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg
for rep in 1 2 3 4; do for mode in normal fifo; do
  if [ $mode = fifo ]; then pre="chrt -f 20"; else pre=""; fi
  echo -n "$mode rep=$rep "; timeout 300 $pre python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gene-index 1 2>&1 | tail -1 | python -c "import sys,json,statistics as st; d=json.loads(sys.stdin.read()); t=d['per_step_total_ms']; print('e2e_ms %.3f' % d['ms_per_step'], 'dev median %.3f mean %.3f max %.3f' % (st.median(t), sum(t)/len(t), max(t)))"
done; done; cat /proc/loadavg
