#!/bin/bash
# round 2, GPU call 16 (1 GPU): trimmed softmax arithmetic in the tcgen05 attention kernel: parity (VL / LLM attention tests, full-size goldens) + prefill time + attention launch list
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c16
timeout -s KILL 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "vl or attention or golden or asr" > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log; tail -n 8 $O.pytest.log | cut -c1-250
timeout -s KILL 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O.launches_prefill.csv python profiles/run_prefill.py 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02c16.launches_prefill.csv')) if len(r) > 10 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:60]; v = float(r[-1].replace(',', ''))
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(14): print(f"{v/1e6:9.3f} ms {100*v/s:5.1f}% x{cnt[k]:5d} {k}")
print("total", s / 1e6, "ms")
PY
