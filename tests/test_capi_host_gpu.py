"""GPU: a host that is not Python runs the sampling stage of the fused path through the C ABI alone.

tests/host/sample_points_host.cpp (plain C++ against include/ndjir_b200.h and libndjir_b200.so: no torch, no Python)
reads parameters and rays from a file, builds the split-fp16 weight planes with ndjir_transpose / ndjir_amax /
ndjir_scale_update / ndjir_pack_h, describes network and sampler as PODs and calls ndjir_sample_points_fwd.  Its
outputs must agree with Engine.sample_points on the same inputs: hit mask and hit count exactly, distances and points
to 1e-5 (the program scales its weight planes by the maximum of the geometric network alone, the engine by the
maximum over all ten networks; the split representation is accurate to 2^-22 either way)."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build_host", "sample_points_host")
EXE_GEO = os.path.join(ROOT, "build_host", "geo_train_host")


def build_host():
    hdr = os.path.join(ROOT, "include", "ndjir_b200.h")      # the PODs live there: a changed header means a rebuild
    for exe in (EXE, EXE_GEO):
        src = os.path.join(ROOT, "tests", "host", os.path.basename(exe) + ".cpp")
        if os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
            continue
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.run(["g++", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", src, "-o", exe,
                        "-L", os.path.join(ROOT, "ndjir_b200"), "-lndjir_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                        "-Wl,-rpath," + os.path.join(ROOT, "ndjir_b200")], check=True)


@pytest.mark.parametrize("kind", ["default", "triplaneline", "no_voxel"])
def test_cpp_host_runs_sample_points(kind, tmp_path):
    from test_engine_gpu import setup, dev
    build_host()
    conf, P, camloc, raydir, color_gt, rnd, eng, model = setup(kind, shape="small")
    r, g = conf.renderer, conf.geometric_network
    B, R, _ = raydir.shape
    passes = 3
    geo = P["geo"]
    v = g.voxel
    grid_kind = {"voxel": 1, "triplaneline": 2}.get(v.type, 0)
    with open(tmp_path / "in.bin", "wb") as f:
        head = [B, R, r.n_samples0, r.n_samples1, r.n_upsamples, r.n_bg_samples,
                {"intersect_with_aabb": 0, "intersect_with_r_sphere": 1}[r.t_near_far_method], g.pe_bands, len(geo) - 1,
                g.skip_layers[0] if len(g.skip_layers) else -1, grid_kind, v.grid_size if grid_kind else 0,
                v.feature_size if grid_kind else 0, passes, 0, 0]
        f.write(struct.pack("<16i", *head))
        f.write(struct.pack("<4f", r.sampling_sigmoid_gain, conf.renderer.bounding_sphere_radius, eng.cskip, 0.0))
        for W, b in geo[:-1]:
            f.write(struct.pack("<2i", *W.shape))
            f.write(np.ascontiguousarray(W, np.float32).tobytes() + np.ascontiguousarray(b, np.float32).tobytes())
        W, b = geo[-1]                                                   # the sdf column of the last layer
        f.write(struct.pack("<2i", W.shape[0], 1))
        f.write(np.ascontiguousarray(W[:, :1], np.float32).tobytes() + np.ascontiguousarray(b[:1], np.float32).tobytes())
        parts = {1: ["voxel"], 2: ["triplane", "triline"]}.get(grid_kind, [])
        for part in parts:
            f.write(eng.params.grid[part].detach().cpu().numpy().astype(np.float32).tobytes())
        for a in (camloc, raydir, rnd["stratified"], rnd["background"]):
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    res = subprocess.run([EXE, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    # the engine on the same inputs, its delayed scales settled over the same number of passes
    args = [dev(camloc), dev(raydir), dev(rnd["stratified"]), dev(rnd["background"])]
    for p in range(passes):
        if p:
            eng.scales.update(eng.stream())
        ms = torch.zeros(1, device="cuda")
        outs = eng.sample_points(*args, mask_sum=ms)
    torch.cuda.synchronize()
    want = [t.cpu().numpy().reshape(-1) for t in outs] + [ms.cpu().numpy()]
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    got, pos = [], 0
    for w in want:
        got.append(raw[pos:pos + w.size]); pos += w.size
    assert pos == raw.size
    names = ("x_fg", "t_fg", "x_bg", "t_bg", "mask", "mask_sum")
    for name, a, b in zip(names, got, want):
        if name in ("mask", "mask_sum"):
            assert np.array_equal(a, b), name
        else:
            err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
            assert err < 1e-5, (name, err)


@pytest.mark.parametrize("kind", ["default", "triplaneline", "no_voxel"])
def test_cpp_host_runs_geometric_network_forward_and_backward(kind, tmp_path):
    """tests/host/geo_train_host.cpp: ndjir_geo_forward, ndjir_geo_normal, the adjoint seed, ndjir_geo_normal_adjoint and
    ndjir_geo_backward from plain C++ for L = 1/2 sum sdf^2 + 0.005 sum feature^2 + 0.05 sum |normal|^2.  sdf, normal,
    every weight / bias gradient (incl. the second-order terms) and the grid-feature gradient must agree with the engine
    running the same passes (1e-5 / 2e-5 of the largest entry: different weight-plane scale, atomic summation order)."""
    from test_engine_gpu import setup
    build_host()
    conf, P, camloc, raydir, color_gt, rnd, eng, model = setup(kind, shape="small")
    g = conf.geometric_network
    geo = P["geo"]
    v = g.voxel
    grid_kind = {"voxel": 1, "triplaneline": 2}.get(v.type, 0)
    rows, passes, Df, nl = 3000, 3, eng.Df, len(geo) - 1
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = (torch.rand((rows, 3), device="cuda", generator=gen) * 1.6 - 0.8).contiguous()
    with open(tmp_path / "in.bin", "wb") as f:
        head = [rows, g.pe_bands, nl, g.skip_layers[0] if len(g.skip_layers) else -1, grid_kind,
                v.grid_size if grid_kind else 0, v.feature_size if grid_kind else 0, passes, Df] + [0] * 7
        f.write(struct.pack("<16i", *head))
        f.write(struct.pack("<1f", eng.cskip))
        W, b = geo[-1]
        for Wl, bl in list(geo[:-1]) + [(W[:, :1], b[:1]), (W[:, 1:], b[1:])]:   # hidden, sdf column, feature block
            f.write(struct.pack("<2i", *Wl.shape))
            f.write(np.ascontiguousarray(Wl, np.float32).tobytes() + np.ascontiguousarray(bl, np.float32).tobytes())
        for part in {1: ["voxel"], 2: ["triplane", "triline"]}.get(grid_kind, []):
            f.write(eng.params.grid[part].detach().cpu().numpy().astype(np.float32).tobytes())
        f.write(x.cpu().numpy().tobytes())
    res = subprocess.run([EXE_GEO, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    # the engine: the same passes (delayed scales settle the same way)
    eng.refresh_transposes()
    for p in range(passes):
        if p:
            eng.scales.update(eng.stream())
        eng.params.zero_grad()
        O = eng.mat("t_O", rows, Df + 6, "fa")
        nrm = eng.buf("t_nrm", rows, 3)
        A, sdf = eng.geo_forward(x, rows, "t", store=True, O=O)
        GZ, Gin = eng.geo_normal(x, rows, "t", A, nrm)
        dsdf = sdf[:rows].reshape(rows, 1).clone()
        dO = eng.mat("t_dO", rows, Df + 6, "fa", grad=True)
        dO.f[:rows, :Df] = 0.01 * O.f[:rows, :Df]
        nbar = (0.1 * nrm[:rows]).contiguous()
        Z2 = eng.geo_normal_adjoint(x, rows, "t", A, GZ, Gin, nbar)
        eng.geo_backward(x, rows, "t", A, dsdf, dO, Z2)
    torch.cuda.synchronize()
    grads = eng.params.export_reference("grad")
    want = [("sdf", sdf[:rows].cpu().numpy().reshape(-1)), ("normal", nrm[:rows].cpu().numpy().reshape(-1))]
    for i in range(nl):
        want += [(f"gW{i}", grads[f"geo.W{i}"].reshape(-1)), (f"gb{i}", grads[f"geo.b{i}"].reshape(-1))]
    Wl, bl = grads[f"geo.W{nl}"], grads[f"geo.b{nl}"]
    want += [("gW_sdf", Wl[:, :1].reshape(-1)), ("gb_sdf", bl[:1]), ("gW_feat", Wl[:, 1:].reshape(-1)), ("gb_feat", bl[1:])]
    if eng.Dg:
        dgrid = eng.mat("geo_dgrid", rows, eng.Dg, "f")
        want.append(("dgrid", dgrid.f[:rows, :eng.Dg].cpu().numpy().reshape(-1)))
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    pos = 0
    for name, w in want:
        a = raw[pos:pos + w.size]
        pos += w.size
        assert np.abs(w).max() > 0, name
        err = float(np.abs(a - w).max() / np.abs(w).max())
        assert err < (1e-5 if name in ("sdf", "normal") else 2e-5), (name, err)
    assert pos == raw.size
