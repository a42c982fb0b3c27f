#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./tools/exp/exp_voxel 2>&1 | tee gpurun_out/exp_voxel.log
# launch list of a full step (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
wc -l gpurun_out/launches_step.csv
# one full capture of the tensor-core product kernel (3 launches from the middle of a step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 3 \
   -o gpurun_out/prof_gemm_tc -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
