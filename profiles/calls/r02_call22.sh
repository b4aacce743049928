#!/bin/bash
# round 2, GPU call 22 (1 GPU): static batching after the GEMV inner-loop rewrite (weights widened once per row, zero-padded rows, 4 / 2 rows per warp)
# and with one CUDA graph per batch composition: the whole batching suite (graph on / off, SIMT twin, samplers, EOS, image + text), then throughput
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c22
timeout -s KILL 300 python -m pytest tests/test_batch_gpu.py -m gpu -q > $O.pytest_batch.log 2>&1; echo "batch rc=$?" | tee -a $O.pytest_batch.log
tail -n 25 $O.pytest_batch.log | cut -c1-300
timeout -s KILL 200 python profiles/run_batch.py 128 64 > $O.run_batch_128.log 2>&1; echo "run_batch rc=$?"
tail -n 10 $O.run_batch_128.log | cut -c1-300
timeout -s KILL 200 python profiles/run_batch.py 1536 32 > $O.run_batch_1536.log 2>&1; echo "run_batch rc=$?"
tail -n 10 $O.run_batch_1536.log | cut -c1-300
