#!/bin/bash
# compute-sanitizer passes over the small-case GPU suite (run through gpurun).  memcheck: every kernel incl. the
# tcgen05 path; racecheck: the shared-memory kernels of the sampler / compositing / grid families (FFMA MLP path).
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_engine_gpu.py tests/test_api_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout=1400 > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log | head -5
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout=1400 -k "ffma or sample_points" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck.log | head -5
