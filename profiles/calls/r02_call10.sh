#!/bin/bash
# round 2, GPU call 10: tcgen05 attention for head_dim 128 (LLM prefill): parity + timing; full GPU suite; N=1 bench; ncu captures of the final tree
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c10
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "prefill or per_layer or attention_implementations" > $O.pytest_attn.log 2>&1; echo "rc=$?" >> $O.pytest_attn.log; tail -n 6 $O.pytest_attn.log
for impl in 0 2; do AHA_ATTN_IMPL=$impl timeout 300 python profiles/run_prefill.py 3 > $O.prefill_attn$impl.log 2>&1; tail -n 2 $O.prefill_attn$impl.log; done
timeout 1200 python -m pytest tests -m gpu -q > $O.pytest_all.log 2>&1; echo "rc=$?" >> $O.pytest_all.log; tail -n 6 $O.pytest_all.log
timeout 600 python bench.py --steps 128 --warmup 8 > $O.bench_vl2_n1.json 2> $O.bench_vl2_n1.err; python -c "import json; d=json.load(open('$O.bench_vl2_n1.json')); print('bench', d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['prefill_secs'], d['cpu_baseline'])"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:decode_step_fused_kernel -s 3 -c 1 -o $O.fused_decode python profiles/run_decode.py 6 > $O.ncu_fused.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O.launches_prefill.csv python profiles/run_prefill.py 1 > $O.ncu_prefill.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:flash_attn_tc_kernel -s 30 -c 1 -o $O.flash_tc128 python profiles/run_prefill.py 1 > $O.ncu_attn128.log 2>&1
