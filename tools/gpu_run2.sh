#!/bin/bash
# engine parity vs the CPU oracle
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout=900 -s --tb=short -rf > gpurun_out/pytest_engine.log 2>&1
grep -v "^  ok " gpurun_out/pytest_engine.log | tail -150
