#!/bin/bash
# round 2, GPU call 12 (1 GPU): persistent 128x256 GEMM (gemm_impl=3) unit tests + shape sweep + prefill A/B; smoke()
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c12
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > $O.pytest_gemm.log 2>&1; echo "rc=$?" >> $O.pytest_gemm.log; tail -n 8 $O.pytest_gemm.log
timeout 300 python profiles/run_gemm.py > $O.gemm_sweep.txt 2>&1; echo "sweep rc=$?"; cat $O.gemm_sweep.txt | tail -n 12
timeout 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2
AHA_GEMM_WIDE=1 timeout 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2
timeout 300 python __graft_entry__.py --smoke > $O.smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 $O.smoke.log
