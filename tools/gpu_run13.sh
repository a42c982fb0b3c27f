#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/pytest_all_gpu.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all_gpu.log | head -40
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err
tail -3 gpurun_out/bench_r1_f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'], d['clocks'])"
