#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=300 --tb=short -rf -s > gpurun_out/pytest_gemm.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gemm.log | head -30
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout=600 --tb=short -rf > gpurun_out/pytest_engine_tc.log 2>&1
grep -E "BAD|passed|failed|FAILED|Error" gpurun_out/pytest_engine_tc.log | head -40
NDJIR_BENCH_DUMP=gpurun_out/gemm_buckets.txt timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err
tail -5 gpurun_out/bench_r1_e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_e.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['ms_per_step_in_kernel'], d['grid_query']['frac'])"
head -30 gpurun_out/gemm_buckets.txt
