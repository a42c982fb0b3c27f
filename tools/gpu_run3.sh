#!/bin/bash
# sampler parity + smoke + first full-size bench + ncu launch list of one step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout=600 --tb=short -rf -k sample_points > gpurun_out/pytest_sampler.log 2>&1
tail -30 gpurun_out/pytest_sampler.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err
tail -5 gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
