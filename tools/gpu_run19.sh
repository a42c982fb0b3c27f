#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_api_gpu.py -m gpu -q --tb=short -rf 2>&1 | tail -8
timeout 900 python tools/bench_grid.py --out gpurun_out/bench_grid.json > gpurun_out/bench_grid.log 2>&1; tail -3 gpurun_out/bench_grid.log | cut -c1-200
for c in triplaneline no_voxel; do
timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_$c.json 2> gpurun_out/bench_r1_$c.err
tail -2 gpurun_out/bench_r1_$c.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_$c.json')); print('$c', d['value'], d['ms_per_step'], d['loss']['loss'])"
done
