#!/bin/bash
# round 2, GPU call 20 (1 GPU): ncu launch list of the bench command restricted to the step kernel (call 19 tried the unfiltered list: ~6000
# launches under ncu do not fit a 7-minute limit).  A decode step IS one launch of decode_step_fused_kernel, so its share of the step is the whole step.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c20
AHA_BENCH_REPS=1 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_step_fused -c 48 --csv --log-file $O.launches_bench_step_kernel.csv \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O.bench_under_ncu.json 2> $O.bench_under_ncu.err; echo "ncu rc=$?"
tail -n 3 $O.bench_under_ncu.err | cut -c1-200
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02c20.launches_bench_step_kernel.csv')) if len(r) > 5 and r[0].isdigit()]
d = [float(r[-1].replace(',', '')) for r in rows]
print('fused launches profiled', len(d), 'unit', rows[0][-2] if rows else None, 'mean', sum(d) / max(len(d), 1), 'min', min(d) if d else None, 'max', max(d) if d else None)
PY
