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


def render_image(pose, intrinsic, resolution, conf, n_rays=None, seed=0, rank=0, world_size=1):
    """Full-frame inference, the loop of the reference's `render_image` (python/renderer.py:212-272): pixel rays are
    generated with `generate_raydir_camloc` (python/helper.py:44-73), rendered in chunks of `n_rays`
    (valid.n_rays = 4000 in default.yaml) with cos_anneal_ratio = 1, and assembled into a (1, 3, H, W) image clipped to
    [0, 1].  With world_size > 1 the chunks are dealt round-robin to the ranks (rays are independent) and the caller
    gathers the partial images (pixels of other ranks are zero).  pose (4,4), intrinsic (3,3) numpy; resolution (W,H)."""
    import numpy as np
    import torch
    from . import scene
    eng = get_engine(conf)
    W, H = resolution
    n_rays = n_rays or 4000
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xy = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1).astype(np.float64)[None]
    raydir, camloc = scene.generate_raydir_camloc(pose[None], intrinsic[None], xy)
    raydir_d = torch.from_numpy(raydir).cuda()
    camloc_d = torch.from_numpy(camloc).cuda()
    out = torch.zeros((H * W, 3), dtype=torch.float32, device="cuda")
    r = conf.renderer
    N = r.n_samples0 + r.n_upsamples * r.n_samples1
    gen = torch.Generator(device="cuda").manual_seed(seed)
    for ci, p0 in enumerate(range(0, H * W, n_rays)):
        if ci % world_size != rank:
            continue
        p1 = min(H * W, p0 + n_rays)
        R = p1 - p0
        u = lambda *shape: torch.rand(shape, device="cuda", generator=gen)
        rnd = dict(stratified=u(1, R, r.n_samples0, 1), background=u(1, R, r.n_bg_samples + 1, 1) * (1 - 1e-5) + 1e-5,
                   diffuse_cdf_the=u(1, R, r.n_thetas), diffuse_cdf_phi=u(1, R, 2 * r.n_thetas),
                   specular_cdf_the=u(1, R, r.n_thetas), specular_cdf_phi=u(1, R, 2 * r.n_thetas),
                   perturb=torch.zeros((1, R, N, 3), device="cuda"))
        gt = torch.zeros((1, R, 3), device="cuda")
        eng.train_step(camloc_d, raydir_d[:, p0:p1].contiguous(), gt, rnd, cos_anneal_ratio=1.0, backward=False,
                       zero_grad=False, keep=True, inference=True)
        out[p0:p1] = eng.debug["color"][:R]
    return out.clamp_(0, 1).reshape(H, W, 3).permute(2, 0, 1)[None]


def sdf_volume(conf, grid_size, batch_size=1 << 20, rank=0, world_size=1):
    """SDF on the `linspace^3` lattice used for marching cubes (python/extract_by_mc.py:47-73 compute_pts_vol): the
    batched geometric_network(x)[0] query, z-slabs dealt to the ranks.  Returns (G_local, G, G) for this rank's x-slabs
    (x is the slowest axis of the reference's meshgrid)."""
    import torch
    eng = get_engine(conf)
    G = grid_size
    rad = eng.rad
    lin = torch.linspace(-rad, rad, G, device="cuda")
    xs = [i for i in range(G) if i % world_size == rank]
    out = torch.empty((len(xs), G, G), dtype=torch.float32, device="cuda")
    yz = torch.stack(torch.meshgrid(lin, lin, indexing="ij"), dim=-1).reshape(-1, 2)
    per = max(1, batch_size // (G * G))
    eng.refresh_transposes()      # W^T and the pre-split lo copies of the weights must match the current parameters
    for s0 in range(0, len(xs), per):
        sl = xs[s0:s0 + per]
        pts = torch.cat([torch.cat([lin[i].expand(G * G, 1), yz], dim=1) for i in sl], dim=0).contiguous()
        sdf = torch.empty((pts.shape[0], 1), dtype=torch.float32, device="cuda")
        eng.geo_forward(pts, pts.shape[0], "smp", store=False, want_feat=False, sdf_out=sdf)
        out[s0:s0 + len(sl)] = sdf.view(len(sl), G, G)
    return out
