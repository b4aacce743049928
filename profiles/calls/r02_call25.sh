#!/bin/bash
# round 2, GPU call 25 (1 GPU, the last of the round): batched GEMV v2 (weights prefetched through a per-thread cp.async ring) -- bit-identity with v1
# (debug_gemm impl 6 vs 5, the batch suite under AHA_BATCH_GEMV=2), then v1 vs v2 throughput for 1..8 lockstep requests
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c25
timeout -s KILL 200 python -m pytest tests/test_batch_gpu.py -m gpu -q > $O.pytest_batch.log 2>&1; echo "batch rc=$?" | tee -a $O.pytest_batch.log
tail -n 25 $O.pytest_batch.log | cut -c1-400
timeout -s KILL 150 python profiles/run_batch.py 128 64 > $O.run_batch_128.log 2>&1; echo "run_batch rc=$?"
tail -n 10 $O.run_batch_128.log | cut -c1-300
