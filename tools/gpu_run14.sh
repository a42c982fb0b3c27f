#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/pytest_all_gpu.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all_gpu.log | head -40
NDJIR_BENCH_DUMP=gpurun_out/gemm_buckets.txt timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err
tail -3 gpurun_out/bench_r1_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_g.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_in_kernel'], d['grid_query'])"
