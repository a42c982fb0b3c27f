#!/bin/bash
mkdir -p gpurun_out
NDJIR_BENCH_DUMP=gpurun_out/gemm_buckets.txt timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err
tail -3 gpurun_out/bench_r1_h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_h.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_in_kernel'], d['roofline']['achieved'])"
head -12 gpurun_out/gemm_buckets.txt
