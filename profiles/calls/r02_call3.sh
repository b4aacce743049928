#!/bin/bash
# round 2, GPU call 3: first hardware run of the packet (LL) version of the fused decode kernel: parity suite + A/B timing against the grid-barrier twin
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c3
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > $O.pytest_parity.log 2>&1; echo "rc=$?" >> $O.pytest_parity.log
tail -n 5 $O.pytest_parity.log
for shape in vl2 q0.6; do
  AHA_SHAPE=$shape timeout 300 python profiles/run_decode.py 64 --sweep "impl=0,tl=1;impl=3,tl=1;impl=0,ctx=512;impl=3,ctx=512" > $O.sweep_$shape.log 2>&1
  grep "tok/s\|FAILED" $O.sweep_$shape.log
done
timeout 900 python -m pytest tests -m gpu -q > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 8 $O.pytest.log
