#!/bin/bash
# round 2, GPU call 15 (1 GPU): pre-split activations (no stand-alone split pass) + CTA-pair GEMM as defaults: parity suites, prefill A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c15
timeout -s KILL 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py -m gpu -q -x > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log; tail -n 12 $O.pytest.log | cut -c1-250
timeout -s KILL 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2
AHA_PRESPLIT=0 timeout -s KILL 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2
AHA_GEMM_PAIR=0 timeout -s KILL 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2
