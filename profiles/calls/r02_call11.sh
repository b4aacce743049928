#!/bin/bash
# round 2, GPU call 11 (8 GPUs): TP test incl. the image-sharded ViT, VL2 replicas + TP-8, config 5 (VL-8B text stack) replicas + TP-8
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c11
timeout 400 python -m pytest tests/test_tp_gpu.py -m gpu -q -s > $O.pytest_tp.log 2>&1; echo "rc=$?" >> $O.pytest_tp.log; tail -n 6 $O.pytest_tp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 64 --warmup 8 > $O.bench_vl2_n8.json 2> $O.bench_vl2_n8.err; echo "vl2 rc=$?"
python -c "import json,sys; d=json.load(open('$O.bench_vl2_n8.json')); print('vl2 N=8 replicas', d['value'], 'tp', d.get('tp'))"
AHA_BENCH_REPS=2 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --preset vl8 --steps 32 --warmup 4 > $O.bench_vl8_n8.json 2> $O.bench_vl8_n8.err; echo "vl8 rc=$?"
tail -n 6 $O.bench_vl8_n8.err
python -c "import json,sys; d=json.load(open('$O.bench_vl8_n8.json')); print('vl8 N=8 replicas', d['value'], d['ms_per_step'], d['config']['prefill_secs'], 'tp', d.get('tp'))"
