#!/bin/bash
# round 2, GPU call 14 (1 GPU): head_dim-72 tower parity; CTA-pair GEMM (gemm_impl=4) tests + sweep (bounded waits: traps instead of hanging)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c14
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "vl or gemm" > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log; tail -n 12 $O.pytest.log
timeout -s KILL 300 python -m pytest tests/test_gemm_pair_gpu.py -m gpu -q -x > $O.pytest_pair.log 2>&1; echo "rc=$?" >> $O.pytest_pair.log; tail -n 25 $O.pytest_pair.log | cut -c1-200
AHA_GEMM_IMPLS=3,4 timeout -s KILL 300 python profiles/run_gemm.py > $O.gemm_pair.txt 2>&1; tail -n 10 $O.gemm_pair.txt | cut -c1-200
nvidia-smi --query-gpu=name,memory.used --format=csv
