#!/bin/bash
# round 2, GPU call 5 (2 GPUs): tensor-parallel fused decode over NVLink packets: parity test + bench with replicas and TP arms; GPU processors test
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c5
nvidia-smi topo -m > $O.topo.txt 2>&1
timeout 600 python -m pytest tests/test_tp_gpu.py tests/test_processors.py -m gpu -q -s > $O.pytest_tp.log 2>&1; echo "rc=$?" >> $O.pytest_tp.log
tail -n 30 $O.pytest_tp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 64 --warmup 8 > $O.bench_n2.json 2> $O.bench_n2.err; echo "bench rc=$?"
tail -n 5 $O.bench_n2.err; cat $O.bench_n2.json
