#!/bin/bash
timeout 900 python -m pytest tests/test_native_gpu.py tests/test_api_gpu.py -m gpu -q --timeout=900 --tb=short -rf 2>&1 | tail -4
timeout 900 python tools/bench_grid.py --only triplane,triline --out gpurun_out/bench_grid2.json > gpurun_out/bench_grid2.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_grid2.log'):
    if l.startswith('{'):
        r=json.loads(l); print(r['kernel'], r['pass_'], r['impl'], r.get('scatter_aggregate'), round(r['ms'],3), round(r['frac_of_hbm_peak'],3))
PY
