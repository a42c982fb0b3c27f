#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r1_2gpu.json 2> gpurun_out/bench_r1_2gpu.err
tail -5 gpurun_out/bench_r1_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['loss']['loss'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference --ref-rays 16 | tail -2 | cut -c1-400
