#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
wc -l gpurun_out/launches_step.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 70 -c 8 \
   -o gpurun_out/prof_gemm_tc -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather4_kernel -s 30 -c 2 \
   -o gpurun_out/prof_gather4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
