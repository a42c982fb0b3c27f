#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=300 --tb=short -rf -s > gpurun_out/pytest_gemm.log 2>&1
grep -E "passed|failed|FAILED|Error|rel err" gpurun_out/pytest_gemm.log | grep -v "tc=0" | head -30
timeout 300 python tools/bench_gemm.py 2>&1 | grep '"tc": 1' | tee gpurun_out/bench_gemm.log
