#!/bin/bash
# round 2, GPU call 19 (1 GPU): the whole GPU suite on the final tree (resampler fallback, bit-exact patchify, video pipeline, prefix cache),
# smoke(), the N=1 bench line, and the ncu launch list of the same bench command (per-launch durations only)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c19
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O.pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O.pytest_gpu.log
tail -n 15 $O.pytest_gpu.log | cut -c1-300
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O.smoke.log | cut -c1-300
timeout -s KILL 420 python bench.py --steps 128 --warmup 8 > $O.bench_vl2_n1.json 2> $O.bench_vl2_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('$O.bench_vl2_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'prefill', d['config']['prefill_secs'], 'launches', d['gpu_launches'])
print('prefix_cache', d.get('prefix_cache'))
print('cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'))
"
timeout -s KILL 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O.launches_bench.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O.bench_under_ncu.json 2> $O.bench_under_ncu.err; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02c19.launches_bench.csv')) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split('(')[0]; agg[name][0] += 1; agg[name][1] += float(r[-1].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print('launches', len(rows), 'total', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'{v[1]/tot:6.1%} {v[0]:6d} {v[1]/v[0]:12.1f} {k[:90]}')
PY
