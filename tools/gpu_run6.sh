#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=300 --tb=short -rf -s > gpurun_out/pytest_gemm.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gemm.log | head -20
timeout 300 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm.log
