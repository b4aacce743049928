#!/bin/bash
# round 2, GPU call 2: anatomy of the grid barrier / activation load inside the fused kernel, with and without the weight stream
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c2
AHA_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k "gqa" > $O.pytest_gqa.log 2>&1; echo "rc=$?" >> $O.pytest_gqa.log
timeout 300 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -s -k "asr06_30s_matches" > $O.pytest_asr_golden.log 2>&1; echo "rc=$?" >> $O.pytest_asr_golden.log
cp aha_b200/libaha_b200.so /tmp/default.so
cp variants/trace.so aha_b200/libaha_b200.so
AHA_SHAPE=vl2 timeout 300 python profiles/run_decode.py 64 --sweep "impl=0,sy=1;impl=0,dbg=512,sy=1;impl=0,dbg=768;impl=0,dbg=2,sy=1;impl=0,dbg=256,sy=1" > $O.anatomy_vl2.log 2>&1
cp /tmp/default.so aha_b200/libaha_b200.so
cat $O.anatomy_vl2.log; tail -n 3 $O.pytest_gqa.log $O.pytest_asr_golden.log
