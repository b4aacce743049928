#!/bin/bash
# round 2, GPU call 4: packet mode with producer flags + one polling warp (A/B vs the barrier twin), device sampler / stream / ASR loop tests
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c4
for shape in vl2 q0.6; do
  AHA_SHAPE=$shape timeout 300 python profiles/run_decode.py 64 --sweep "impl=0,tl=1;impl=3" > $O.sweep_$shape.log 2>&1
  grep "tok/s\|FAILED" $O.sweep_$shape.log
done
timeout 600 python -m pytest tests/test_sampling_gpu.py -m gpu -q > $O.pytest_sampling.log 2>&1; echo "rc=$?" >> $O.pytest_sampling.log
tail -n 25 $O.pytest_sampling.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_sampling_gpu.py > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 8 $O.pytest.log
