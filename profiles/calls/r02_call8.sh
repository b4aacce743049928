#!/bin/bash
# round 2, GPU call 8: packet modes with gpu-scope ops on one GPU and the peer pointers in a device table: A/B against the barrier kernel
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c8
AHA_SHAPE=vl2 timeout 300 python profiles/run_decode.py 64 --sweep "impl=3;impl=4,tl=1;impl=2,tl=1" > $O.sweep_vl2.log 2>&1
grep "tok/s\|FAILED" $O.sweep_vl2.log
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "implementations_agree or production_row or gqa" > $O.pytest.log 2>&1; tail -n 3 $O.pytest.log
