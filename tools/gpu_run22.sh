#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf 2>&1 | tail -4
timeout 900 python tools/bench_grid.py --out gpurun_out/bench_grid.json > gpurun_out/bench_grid.log 2>&1; tail -1 gpurun_out/bench_grid.log | cut -c1-100
