#!/bin/bash
# round 2, GPU call 17 (8 GPUs): config 5 (Qwen3-VL-8B shape, real head_dim-72 tower, 4 x 2048^2 images) replicas + TP-8; then VL2 replicas + TP-8; TP-2 test with the image-sharded ViT
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c17
AHA_BENCH_REPS=1 timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --preset vl8 --steps 32 --warmup 4 > $O.bench_vl8_n8.json 2> $O.bench_vl8_n8.err; echo "vl8 rc=$?"
tail -n 4 $O.bench_vl8_n8.err | cut -c1-300
python -c "import json,sys; d=json.load(open('$O.bench_vl8_n8.json')); print('vl8 N=8 replicas', d['value'], d['ms_per_step'], d['config']['prefill_secs'], 'tp', d.get('tp'))"
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 64 --warmup 8 > $O.bench_vl2_n8.json 2> $O.bench_vl2_n8.err; echo "vl2 rc=$?"
python -c "import json,sys; d=json.load(open('$O.bench_vl2_n8.json')); print('vl2 N=8 replicas', d['value'], 'tp', d.get('tp'))"
timeout -s KILL 300 python -m pytest tests/test_tp_gpu.py -m gpu -q -s > $O.pytest_tp.log 2>&1; echo "rc=$?" >> $O.pytest_tp.log; tail -n 6 $O.pytest_tp.log | cut -c1-300
