#!/bin/bash
# round 2, GPU call 9 (4 GPUs): TP test, VL2 replicas + TP-4, config 5 (Qwen3-VL-8B text stack, 4 x 2048^2 images + 512 ids) replicas (1-GPU point) + TP-4
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c9
timeout 400 python -m pytest tests/test_tp_gpu.py -m gpu -q -s > $O.pytest_tp.log 2>&1; echo "rc=$?" >> $O.pytest_tp.log; tail -n 4 $O.pytest_tp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 64 --warmup 8 > $O.bench_vl2_n4.json 2> $O.bench_vl2_n4.err; echo "vl2 rc=$?"
python -c "import json,sys; d=json.load(open('$O.bench_vl2_n4.json')); print('vl2 N=4 replicas', d['value'], 'tp', d.get('tp'))"
AHA_BENCH_REPS=2 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --preset vl8 --steps 32 --warmup 4 > $O.bench_vl8_n4.json 2> $O.bench_vl8_n4.err; echo "vl8 rc=$?"
tail -n 12 $O.bench_vl8_n4.err
python -c "import json,sys; d=json.load(open('$O.bench_vl8_n4.json')); print('vl8 N=4 replicas', d['value'], d['ms_per_step'], d['decode_impl'], d['config']['prefill_secs'], 'tp', d.get('tp'))"
