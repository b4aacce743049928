#!/bin/bash
# round 2, GPU call 23 (1 GPU): the N=1 bench line on the final tree (prefix_cache and batch records measured last on the handle) and the
# request-level suites that share code with the batch entry (prefix cache, sampling / ASR loop, generate / error / clear_cache parity tests)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c23
timeout -s KILL 420 python bench.py --gpus 1 --steps 128 --warmup 8 > $O.bench_vl2_n1.json 2> $O.bench_vl2_n1.err; echo "bench rc=$?"
tail -n 3 $O.bench_vl2_n1.err | cut -c1-300
python -c "
import json
d=json.load(open('$O.bench_vl2_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'prefill', d['config']['prefill_secs'], 'launches', d['gpu_launches'])
print('prefix_cache', d.get('prefix_cache'))
print('batch', d.get('batch'))
print('cpu', (d.get('cpu_baseline') or {}).get('value'))
"
timeout -s KILL 300 python -m pytest tests/test_prefix_cache_gpu.py tests/test_sampling_gpu.py -m gpu -q > $O.pytest_requests.log 2>&1; echo "requests rc=$?" | tee -a $O.pytest_requests.log
tail -n 5 $O.pytest_requests.log | cut -c1-300
timeout -s KILL 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "greedy or vl_prefill or asr_prefill or error or clear_cache or video or text_only" > $O.pytest_parity_subset.log 2>&1; echo "parity rc=$?" | tee -a $O.pytest_parity_subset.log
tail -n 5 $O.pytest_parity_subset.log | cut -c1-300
