#!/bin/bash
mkdir -p gpurun_out
NDJIR_BENCH_DUMP=gpurun_out/gemm_buckets.txt timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err
tail -5 gpurun_out/bench_r1_c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['ms_per_step_in_kernel'])"
head -70 gpurun_out/gemm_buckets.txt
