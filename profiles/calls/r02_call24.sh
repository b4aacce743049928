#!/bin/bash
# round 2, GPU call 24 (1 GPU): continuous batching (batch_open / _add / _step / _close) + the static batching suite again (its prefill now goes
# through the shared helper)
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c24
timeout -s KILL 240 python -m pytest tests/test_batch_gpu.py -m gpu -q > $O.pytest_batch.log 2>&1; echo "batch rc=$?" | tee -a $O.pytest_batch.log
tail -n 30 $O.pytest_batch.log | cut -c1-400
