#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short -rf > gpurun_out/pytest_all_gpu.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_all_gpu.log | head -20
timeout 900 python tools/bench_grid.py --only voxel,triplane,triline --out gpurun_out/bench_grid2.json > gpurun_out/bench_grid2.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_grid2.log'):
    if l.startswith('{'):
        r=json.loads(l); print(r['kernel'], r['pass_'], r['impl'], r.get('scatter_aggregate'), round(r['ms'],3), round(r['frac_of_hbm_peak'],3))
PY
