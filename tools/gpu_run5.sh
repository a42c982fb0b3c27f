#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout=600 --tb=short -rf > gpurun_out/pytest_engine_tc.log 2>&1
grep -E "BAD|passed|failed|FAILED|Error" gpurun_out/pytest_engine_tc.log | head -40
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err
tail -5 gpurun_out/bench_r1_b.err; cat gpurun_out/bench_r1_b.json
