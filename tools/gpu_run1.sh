#!/bin/bash
# first GPU pass: parity tests, grid micro-bench, ncu launch list + one full capture of the voxel kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_grid.py --out gpurun_out/bench_grid.json > gpurun_out/bench_grid.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_grid.csv \
   python tools/bench_grid.py --only voxel --no-ref --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|scatter_kernel' -s 6 -c 4 \
   -o gpurun_out/prof_voxel -f python tools/bench_grid.py --only voxel --no-ref --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
tail -30 gpurun_out/bench_grid.log
