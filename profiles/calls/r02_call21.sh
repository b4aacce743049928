#!/bin/bash
# round 2, GPU call 21 (1 GPU): first hardware run of static batching (batched GEMV, batched decode attention, generate_batch) + the per-op
# decode tests that cover the refactored decode attention body + aggregate throughput of 1..8 lockstep requests on the VL2 text stack
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c21
timeout -s KILL 300 python -m pytest tests/test_batch_gpu.py -m gpu -q -x > $O.pytest_batch.log 2>&1; echo "batch rc=$?" | tee -a $O.pytest_batch.log
tail -n 30 $O.pytest_batch.log | cut -c1-300
timeout -s KILL 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "decode_implementations_agree or long_context_decode or other_gqa or teacher_forced" > $O.pytest_perop.log 2>&1; echo "perop rc=$?" | tee -a $O.pytest_perop.log
tail -n 4 $O.pytest_perop.log | cut -c1-300
timeout -s KILL 240 python profiles/run_batch.py 128 64 > $O.run_batch.log 2>&1; echo "run_batch rc=$?"
tail -n 12 $O.run_batch.log | cut -c1-300
