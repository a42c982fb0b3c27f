#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_multi_gpu.py 2>&1 | grep -E "MULTI_GPU|Error|error" | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1_2gpu.json 2> gpurun_out/bench_r1_2gpu.err
tail -3 gpurun_out/bench_r1_2gpu.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss']['loss'])"
