#!/bin/bash
# round 2, GPU call 6: tcgen05 flash attention (parity, prefill timing, ncu), ld.acquire barrier A/B, bench lines for configs 2 and 4
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c6
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "vl or asr" > $O.pytest_vl_asr.log 2>&1; echo "rc=$?" >> $O.pytest_vl_asr.log
tail -n 15 $O.pytest_vl_asr.log
for impl in 0 2; do AHA_ATTN_IMPL=$impl timeout 300 python profiles/run_prefill.py 4 > $O.prefill_attn$impl.log 2>&1; tail -n 3 $O.prefill_attn$impl.log; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"flash_attn|split_qkv" -c 120 --csv --log-file $O.attn_launches.csv python profiles/run_prefill.py 1 > $O.ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:flash_attn_tc_kernel -s 6 -c 1 -o $O.flash_tc python profiles/run_prefill.py 1 > $O.ncu_attn_full.log 2>&1
timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s -k "golden" > $O.pytest_golden.log 2>&1; echo "rc=$?" >> $O.pytest_golden.log
grep -h "max |dlogit|\|passed\|failed" $O.pytest_golden.log
AHA_SHAPE=vl2 timeout 200 python profiles/run_decode.py 64 --sweep "impl=3;impl=3" > $O.ldacq_base.log 2>&1
cp aha_b200/libaha_b200.so /tmp/default.so; cp variants/ldacq.so aha_b200/libaha_b200.so
AHA_SHAPE=vl2 timeout 200 python profiles/run_decode.py 64 --sweep "impl=3;impl=3" > $O.ldacq_var.log 2>&1
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "qwen3_teacher or implementations_agree" > $O.ldacq_pytest.log 2>&1
cp /tmp/default.so aha_b200/libaha_b200.so
grep -h "tok/s" $O.ldacq_base.log $O.ldacq_var.log; tail -n 2 $O.ldacq_pytest.log
for p in q0.6 asr0.6; do timeout 400 python bench.py --preset $p --steps 64 --warmup 8 > $O.bench_$p.json 2> $O.bench_$p.err; cat $O.bench_$p.json | cut -c1-600; done
