"""`pb_render` with the reference's outputs (python/renderer.py:32-209), forward only (the backward runs inside
loss.total_loss / Engine.train_step)."""
from .engine import get_engine


def pb_render(camloc, raydir, color_gt, rnd, cos_anneal_ratio, conf):
    """Runs sample_points + pb_render and returns the reference's dictionary entries that exist as buffers:
    color_pixel (B,R,3), sdf_x_fg, grad_x_fg, alpha_fg, trans_fg, weights_fg, weights_bg, normal_pixel, ..."""
    eng = get_engine(conf)
    eng.train_step(camloc, raydir, color_gt, rnd, cos_anneal_ratio=cos_anneal_ratio, backward=False, keep=True)
    d = eng.debug
    B, R, N, Nb, M = d["dims"]
    NR, P, Df = B * R, B * R * N, eng.Df
    att = d["ATT"][:P]
    return dict(
        color_pixel=d["color"][:NR].view(B, R, 3), sdf_x_fg=d["sdf"][:P].view(B, R, N, 1),
        grad_x_fg=d["nrm"][:P].view(B, R, N, 3), alpha_fg=d["alpha_fg"][:P].view(B, R, N, 1),
        trans_fg=d["T"][:NR, :N].reshape(B, R, N, 1), weights_fg=d["w"][:NR, :N].reshape(B, R, N, 1),
        weights_bg=d["w"][:NR, N:].reshape(B, R, Nb, 1), normal_pixel=d["nhat"][:NR].view(B, R, 3),
        feature=d["O"][:P, :Df].reshape(B, R, N, Df), implicit=att[:, 0].reshape(B, R, N, 1),
        roughness=att[:, 1].reshape(B, R, N, 1), specular_reflectance=att[:, 2:5].reshape(B, R, N, 3),
        photogrammetric=att[:, 5].reshape(B, R, N, 1), x_fg=d["x_fg"].view(B, R, N, 3), mask=d["mask"])


def render_image(pose, intrinsic, resolution, conf, n_rays=None, seed=0, rank=0, world_size=1, process_group=None):
    """Full-frame inference, the loop of the reference's `render_image` (python/renderer.py:212-272) with cos_anneal_ratio
    = 1, as DEVICE work: pixel rays (helper.generate_raydir_camloc, python/helper.py:44-73), the random inputs of
    sample_points / pb_render, the forward path and the store into the image are captured ONCE in a CUDA graph that reads
    a device chunk counter, and the graph is replayed per chunk of `n_rays` (valid.n_rays = 4000 in default.yaml) - no
    host round trip per chunk (the reference copies rays in and colours out 480 times per 1600 x 1200 frame).
    With world_size > 1 the chunks are dealt round-robin to the ranks and the partial images are summed onto rank 0
    (every pixel belongs to exactly one rank).  pose (4,4), intrinsic (3,3) numpy; resolution (W,H).
    Returns the (1, 3, H, W) image clipped to [0, 1] (complete on rank 0)."""
    import numpy as np
    import torch
    from . import _lib
    eng = get_engine(conf)
    W, H = resolution
    n = int(n_rays or 4000)
    n_pix = H * W
    n_chunks = (n_pix + n - 1) // n
    r = conf.renderer
    N = r.n_samples0 + r.n_upsamples * r.n_samples1
    dev = eng.device
    st = lambda: torch.cuda.current_stream().cuda_stream    # noqa: E731
    kinv = torch.from_numpy(np.linalg.inv(np.asarray(intrinsic, np.float64)).reshape(9).copy()).to(dev)
    rot = torch.from_numpy(np.ascontiguousarray(np.asarray(pose, np.float64)[:3, :3]).reshape(9).copy()).to(dev)
    camloc = torch.from_numpy(np.asarray(pose, np.float64)[:3, 3].astype(np.float32).reshape(1, 3)).to(dev)
    key = ("render_image", n, W, H, world_size, seed)
    state = eng._graphs.get(key)
    if state is None:
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)    # noqa: E731
        bufs = dict(raydir=f(1, n, 3), gt=torch.zeros((1, n, 3), device=dev), image=f(n_pix, 3),
                    chunk=torch.zeros(1, dtype=torch.int32, device=dev), kinv=kinv.clone(), rot=rot.clone(),
                    camloc=camloc.clone(), seed=seed,
                    rnd=dict(stratified=f(1, n, r.n_samples0, 1), background=f(1, n, r.n_bg_samples + 1, 1),
                             diffuse_cdf_the=f(1, n, r.n_thetas), diffuse_cdf_phi=f(1, n, 2 * r.n_thetas),
                             specular_cdf_the=f(1, n, r.n_thetas), specular_cdf_phi=f(1, n, 2 * r.n_thetas),
                             perturb=torch.zeros((1, n, N, 3), device=dev)))

        def one_chunk(b=bufs):
            ck = b["chunk"]
            _lib.call("ndjir_generate_rays", n, W, n_pix, 0, ck, b["kinv"], b["rot"], b["raydir"], st())
            lows = dict(background=1e-5)
            for i, (k, t) in enumerate(sorted(b["rnd"].items())):
                if k != "perturb":
                    _lib.call("ndjir_uniform", t.numel(), lows.get(k, 0.0), 1.0, b["seed"] * 16 + i, ck, t, st())
            eng.train_step(b["camloc"], b["raydir"], b["gt"], b["rnd"], cos_anneal_ratio=1.0, backward=False,
                           zero_grad=False, keep=True, inference=True)
            _lib.call("ndjir_store_chunk", n, n_pix, 0, ck, eng.debug["color"], b["image"], st())
            _lib.call("ndjir_counter_add", ck, b["stride"], st())
        bufs["stride"] = world_size
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # eager warm-up: sizes every scratch buffer, seeds the split-fp16 scales
            one_chunk()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_chunk()
        state = eng._graphs[key] = (graph, bufs)
    graph, b = state
    b["kinv"].copy_(kinv); b["rot"].copy_(rot); b["camloc"].copy_(camloc)
    b["image"].zero_()
    b["chunk"].fill_(rank)
    for _ in range(rank, n_chunks, world_size):
        graph.replay()
    img = b["image"]
    if world_size > 1:
        import torch.distributed as dist
        if dist.is_initialized():      # (a caller that only simulates the dealing keeps the partial images)
            dist.reduce(img, dst=0, op=dist.ReduceOp.SUM, group=process_group)
    return img.reshape(H, W, 3).permute(2, 0, 1)[None].clone()


def sdf_volume(conf, grid_size, batch_size=1 << 20, rank=0, world_size=1, process_group=None, gather=False):
    """SDF on the `linspace^3` lattice used for marching cubes (python/extract_by_mc.py:47-73 compute_pts_vol): the
    batched geometric_network(x)[0] query with x-planes dealt round-robin to the ranks; lattice points are generated on
    the device (ndjir_lattice_points).  Returns (G_local, G, G) for this rank's x-planes (x is the slowest axis of the
    reference's meshgrid); with gather=True every rank contributes to the full (G, G, G) volume summed onto rank 0."""
    import torch
    from . import _lib
    eng = get_engine(conf)
    G = grid_size
    rad = eng.rad
    xs = [i for i in range(G) if i % world_size == rank]
    out = torch.empty((len(xs), G, G), dtype=torch.float32, device="cuda")
    per = max(1, batch_size // (G * G))
    pts = torch.empty((per * G * G, 3), dtype=torch.float32, device="cuda")
    eng.refresh_transposes()      # W^T and the split copies of the weights must match the current parameters
    if eng.fused_calls and len(xs):
        # the whole extraction is ONE C-ABI call (ndjir_sdf_lattice, csrc/fused_path.cu): lattice points and network
        # evaluation are sequenced inside the library, on the engine's own scratch buffers and scale slots
        from . import h16
        rows = per * G * G
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
        _lib.call("ndjir_sdf_lattice", eng.geo_net_desc(), G, xs[0], world_size, len(xs), rad, rows, pts, ws, out,
                  torch.cuda.current_stream().cuda_stream)
        xs = []
    for s0 in range(0, len(xs), per):
        cnt = min(per, len(xs) - s0)
        npts = cnt * G * G
        _lib.call("ndjir_lattice_points", npts, G, xs[s0], world_size, rad, pts, torch.cuda.current_stream().cuda_stream)
        sdf = out[s0:s0 + cnt].view(npts, 1)
        eng.geo_forward(pts, npts, "smp", store=False, want_feat=False, sdf_out=sdf)
    if gather:
        import torch.distributed as dist
        full = torch.zeros((G, G, G), dtype=torch.float32, device="cuda")
        full[rank::world_size] = out
        if world_size > 1 and dist.is_initialized():
            dist.reduce(full, dst=0, op=dist.ReduceOp.SUM, group=process_group)
        return full
    return out
