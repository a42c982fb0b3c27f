#!/bin/bash
# One pass on a B200 box (run through gpurun): GPU test suite, smoke, benchmark, launch list, grid micro-benchmark.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/pytest_all_gpu.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all_gpu.log | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
NDJIR_BENCH_DUMP=gpurun_out/gemm_buckets.txt timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['grid_query']['frac'], d['optimizer_step']['frac'], d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log | cut -c1-300
timeout 900 python tools/bench_grid.py > gpurun_out/bench_grid.log 2>&1; tail -3 gpurun_out/bench_grid.log | cut -c1-400
