#!/bin/bash
# round 2, GPU call 7: hybrid fused kernel (barriers inside the layer, packets for the residual exchanges) A/B; 2-CTA GEMM variant; bf16 checkpoint test
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c7
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "vl or asr" > $O.pytest_vl_asr.log 2>&1; echo "rc=$?" >> $O.pytest_vl_asr.log
tail -n 4 $O.pytest_vl_asr.log
for shape in vl2 q0.6; do
  AHA_SHAPE=$shape timeout 300 python profiles/run_decode.py 64 --sweep "impl=3;impl=4,tl=1;impl=2;impl=4,ctx=512;impl=3,ctx=512" > $O.sweep_$shape.log 2>&1
  grep "tok/s\|FAILED" $O.sweep_$shape.log
done
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "qwen3 or gqa or bf16" > $O.pytest_parity.log 2>&1; echo "rc=$?" >> $O.pytest_parity.log
tail -n 6 $O.pytest_parity.log
timeout 200 python -m pytest tests/test_gemm_gpu.py -m gpu -q -s -k timing > $O.gemm_base.log 2>&1; grep "GEMM" $O.gemm_base.log
AHA_ATTN_IMPL=0 timeout 200 python profiles/run_prefill.py 3 > $O.prefill_base.log 2>&1; tail -n 2 $O.prefill_base.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"flash_attn|split_qkv" -c 60 --csv --log-file $O.attn_launches.csv python profiles/run_prefill.py 1 > $O.ncu_attn.log 2>&1
timeout 400 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s -k "vl2_1080p_matches" > $O.pytest_golden.log 2>&1; grep -h "max |dlogit|\|passed\|failed" $O.pytest_golden.log
cp aha_b200/libaha_b200.so /tmp/default.so; cp variants/gemm2cta.so aha_b200/libaha_b200.so
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -s > $O.gemm_2cta.log 2>&1; grep "GEMM\|passed\|failed" $O.gemm_2cta.log
AHA_ATTN_IMPL=0 timeout 200 python profiles/run_prefill.py 3 > $O.prefill_2cta.log 2>&1; tail -n 2 $O.prefill_2cta.log
cp /tmp/default.so aha_b200/libaha_b200.so
