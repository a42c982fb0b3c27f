"""ndjir_geo_sdf_forward at default.yaml widths: layer-by-layer products vs the one-kernel on-chip chain (mlp_h_chain)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ndjir_b200 import _lib, scene, h16
from ndjir_b200.config import make_conf
from ndjir_b200.engine import Engine
from test_geo_chain_gpu import sdf_through_c_abi

for kind in ("default", "no_voxel"):
    conf = make_conf(kind)
    eng = Engine(conf)
    eng.params.load_reference(scene.init_params(conf, seed=313))
    eng.params.init_grid_on_device(scene.grid_shapes(conf), std=1e-3, seed=313)
    eng.refresh_transposes()
    for rows in (32768, 131072, 1 << 20):
        pts = (torch.rand((rows, 3), device="cuda") * 1.6 - 0.8).contiguous()
        for chain in (0, 1):
            _lib.call("ndjir_set_option", "mlp_h_chain", chain)
            sdf_through_c_abi(eng, pts)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sdf_through_c_abi_once = sdf_through_c_abi   # (3 passes inside)
                sdf_through_c_abi(eng, pts)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 3)
            print(f"{kind:9s} rows {rows:8d} chain={chain}: {np.median(ts):.3f} ms per evaluation "
                  f"({rows / np.median(ts) / 1e3:.1f} M points/s)", flush=True)
        _lib.call("ndjir_set_option", "mlp_h_chain", 0)
