#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_r1_${n}gpu.json 2> gpurun_out/bench_r1_${n}gpu.err
tail -3 gpurun_out/bench_r1_${n}gpu.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_${n}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss']['loss'], d['clocks'])"
done
