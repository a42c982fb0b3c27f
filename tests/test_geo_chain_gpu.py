"""GPU: the SDF network as ONE kernel with the activations on chip (csrc/gemm_h_chain.cu, reached through
ndjir_geo_sdf_forward) against the oracle's geometric_network in float64 and against the layer-by-layer evaluation of
the same entry point, at the full widths of default.yaml (256-wide layers, the 213 + 43 skip layer, voxel / triplane +
triline / no grid features), for row counts that are not multiples of the 128-row tile."""
import numpy as np
import pytest
import torch

from ndjir_b200 import _lib, h16
from oracle import cpu_render as CR

pytestmark = pytest.mark.gpu


def sdf_through_c_abi(eng, pts):
    rows = pts.shape[0]
    eng._reserve = rows
    A0 = eng.mat("smp_A0", rows, eng.din, "fa")
    widest = max(L.K for L in eng.params.nets["geo"][1:])
    act = [eng.mat(f"geo_pp{i}", rows, widest, "a") for i in (0, 1)]
    gw = sum(w for _, w, _ in eng._grid_parts())
    gtmp = eng.buf("gq_fused", rows, max(gw, 1)) if gw else None
    eng._reserve = 0
    ws = h16.GeoScratch()
    ws.enc, ws.ld_enc = A0.f.data_ptr(), eng.ld0
    ws.grid_tmp = gtmp.data_ptr() if gtmp is not None else None
    ws.ench = A0.hmat(0)
    ws.act[0], ws.act[1] = act[0].hmat(0), act[1].hmat(0)
    out = torch.empty(rows, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for i in range(3):        # the delayed scales settle over the first passes
        if i:
            eng.scales.update(st)
        _lib.call("ndjir_geo_sdf_forward", eng.geo_net_desc(), rows, pts, out, ws, st)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("kind,rows", [("default", 1000), ("triplaneline", 4096 + 77), ("no_voxel", 128), ("default", 40000)])
def test_chained_sdf_network_matches_oracle_and_layerwise(kind, rows):
    from test_engine_gpu import setup
    conf, P, camloc, raydir, color_gt, rnd, eng, model = setup(kind, shape="full")
    eng.refresh_transposes()
    g = torch.Generator(device="cuda").manual_seed(3)
    pts = (torch.rand((rows, 3), device="cuda", generator=g) * 1.6 - 0.8).contiguous()
    with torch.no_grad():
        want = model.geometric_network(pts.cpu().double())[0].reshape(-1).numpy()
    got_layers = sdf_through_c_abi(eng, pts).cpu().numpy()
    _lib.call("ndjir_set_option", "mlp_h_chain", 1)       # off by default (measured slower)
    try:
        got_chain = sdf_through_c_abi(eng, pts).cpu().numpy()
    finally:
        _lib.call("ndjir_set_option", "mlp_h_chain", 0)
    scale = np.abs(want).max()
    e_chain, e_layers = np.abs(got_chain - want).max() / scale, np.abs(got_layers - want).max() / scale
    assert e_layers < 1e-5, e_layers
    assert e_chain < 1e-5, e_chain                      # the forward bar of BASELINE.json
    assert np.abs(got_chain - got_layers).max() / scale < 1e-5
