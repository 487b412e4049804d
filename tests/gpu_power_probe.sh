#!/bin/bash
# per-kernel times + power/clock samples for two timesteps, with and without host polling inside the CG solve
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active --format=csv,noheader -lms 50 > gpurun_out/power_samples.csv &
SMI=$!
for cfg in "0.625 " "0.0625 " "0.0625 RXB_QEQ_WAIT=1" "0.625 RXB_QEQ_WAIT=1"; do
  set -- $cfg
  echo "== dt $1 $2 $(date +%s.%N)" >> gpurun_out/power_marks.txt
  env $2 python tests/gpu_perf_probe.py 8 40 $1 1 2>&1 | tail -1
done
echo "== end $(date +%s.%N)" >> gpurun_out/power_marks.txt
kill $SMI
python - <<'PY'
import collections
rows=[l.strip().split(', ') for l in open('gpurun_out/power_samples.csv') if l.strip()]
clk=[int(r[0].split()[0]) for r in rows]; pw=[float(r[1].split()[0]) for r in rows]
import statistics
print('samples',len(rows),'clock min/median/max',min(clk),statistics.median(clk),max(clk),'power median/max',statistics.median(pw),max(pw))
print('sw_power_cap active samples', sum(1 for r in rows if 'Active' in r[2] and 'Not' not in r[2]))
lo=[(c,p) for c,p in zip(clk,pw) if c<1900]
print('samples below 1900 MHz:',len(lo), lo[:10])
PY
