#!/bin/bash
# compute-sanitizer passes over the small-case GPU suite (run through gpurun).  memcheck: every kernel incl. the
# tcgen05 path; racecheck: the shared-memory kernels of the sampler / compositing / grid families (FFMA MLP path).
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_engine_gpu.py tests/test_api_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout=1400 > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck.log | head -5
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout=1400 -k "ffma or sample_points" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck.log | head -5
# later round-1 kernels: brick-ordered sweeps, cosine / Lanczos / hash families, fused bias gradients, optimizer
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_native_gpu.py tests/test_gemm_gpu.py tests/test_solver.py -m gpu -q --timeout=1100 -k "binned or lanczos or hash or cosine or triplane or wgrad_bias or tv_loss or fused_step or presplit" > gpurun_out/sanitizer_memcheck2.log 2>&1
echo "memcheck (new kernels) exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck2.log | head -5
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_native_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout=1100 -k "binned or coarse or sample_points" > gpurun_out/sanitizer_racecheck2.log 2>&1
echo "racecheck (new kernels) exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck2.log | head -5
