#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=300 --tb=short -rf -s > gpurun_out/pytest_gemm.log 2>&1
grep -E "gemm M=|raw hi|passed|failed|FAILED|Error|error" gpurun_out/pytest_gemm.log | head -60
