#!/bin/bash
# round 2, GPU call 1: full GPU suite, first hardware run of the decode_impl 3/4/5 variants, streaming-rate switches, per-stage trace
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O.smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "not full_size_oracle_golden or q06" > $O.pytest.log 2>&1; echo "pytest rc=$?" >> $O.pytest.log
AHA_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_experimental_gpu.py -m gpu -q > $O.pytest_exp.log 2>&1; echo "pytest rc=$?" >> $O.pytest_exp.log
for shape in vl2 q0.6; do
  AHA_SHAPE=$shape timeout 300 python profiles/run_decode.py 64 --sweep "impl=0,tl=1;impl=0,dbg=1;impl=0,dbg=2;impl=0,dbg=3;impl=0,dbg=259;impl=0,dbg=258;impl=0,dbg=256;impl=0,stages=8;impl=1" > $O.sweep_$shape.log 2>&1
  for i in 3 4 5; do AHA_SHAPE=$shape timeout 200 python profiles/run_decode.py 64 --sweep "impl=$i,tl=1" > $O.impl${i}_$shape.log 2>&1; done
done
cp aha_b200/libaha_b200.so /tmp/default.so
cp variants/trace.so aha_b200/libaha_b200.so
AHA_SHAPE=vl2 timeout 300 python profiles/run_decode.py 64 --sweep "impl=0,st=1;impl=0,dbg=2,st=1;impl=0,dbg=3,st=1" > $O.trace_vl2.log 2>&1
AHA_SHAPE=q0.6 timeout 300 python profiles/run_decode.py 64 --sweep "impl=0,st=1" > $O.trace_q06.log 2>&1
cp /tmp/default.so aha_b200/libaha_b200.so
tail -3 $O.pytest.log $O.pytest_exp.log; grep -h "tok/s" $O.sweep_*.log $O.impl*.log $O.trace_*.log
