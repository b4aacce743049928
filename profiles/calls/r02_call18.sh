#!/bin/bash
# round 2, GPU call 18 (1 GPU): first hardware run of the prefix cache / prefill continuation / video processor tests, then the sampling / ASR-loop suite (generate_impl changed), the fp16-KV accuracy measurement (variants/kv16.so = the library built with -DAHA_KV_ROUND_FP16), and the N=1 bench with the
# prefix_cache record
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c18
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O.gpu.txt 2>&1
timeout -s KILL 420 python -m pytest tests/test_prefix_cache_gpu.py tests/test_processors.py -m gpu -q -s > $O.pytest_new.log 2>&1; echo "new rc=$?" | tee -a $O.pytest_new.log
tail -n 25 $O.pytest_new.log | cut -c1-400
timeout -s KILL 300 python -m pytest tests/test_sampling_gpu.py -m gpu -q > $O.pytest_sampling.log 2>&1; echo "sampling rc=$?" | tee -a $O.pytest_sampling.log
tail -n 6 $O.pytest_sampling.log | cut -c1-400
cp aha_b200/libaha_b200.so /tmp/default.so; cp variants/kv16.so aha_b200/libaha_b200.so
timeout -s KILL 300 python profiles/run_kv16.py > $O.kv16.log 2>&1; echo "kv16 rc=$?" | tee -a $O.kv16.log
cp /tmp/default.so aha_b200/libaha_b200.so
tail -n 5 $O.kv16.log | cut -c1-600
timeout -s KILL 420 python bench.py --steps 128 --warmup 8 --no-cpu-baseline > $O.bench_vl2_n1.json 2> $O.bench_vl2_n1.err; echo "bench rc=$?"
tail -n 3 $O.bench_vl2_n1.err | cut -c1-300
python -c "
import json
d=json.load(open('$O.bench_vl2_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'prefill', d['config']['prefill_secs'])
print('prefix_cache', d.get('prefix_cache'))
print('cpu', d.get('cpu_baseline', {}).get('value'))
"
