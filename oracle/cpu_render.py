"""CPU restatement (torch autograd, float64 or float32) of the nnabla-composed part of NDJIR's hot path:
sample placement, the SDF / colour / illumination MLPs, NeuS alpha compositing, PBR shading and the losses.

TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.py header): imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by ndjir_b200/.

PINNED on the reference's own Python: tests/golden/make_golden.py::golden_render EXECUTES python/sampler.py,
network.py, renderer.py, specular_brdf.py and loss.py where they lie under /root/reference - through a torch-backed
stand-in for the slice of nnabla they use (tests/golden/nnabla_standin.py; nnabla 1.29 is not installable here) - on
the seeded cases of tests/golden/cases.py (the real per-ray shapes: 64 + 4 x 16 samples, 32 background samples, 128
light directions; default widths incl. the 213 + 43 skip layer) and stores sample_points / pb_render / total_loss
outputs and every parameter gradient in tests/golden/render_*.npz; tests/test_render_golden.py holds this file to
those vectors at 1e-9 (float64 on both sides).  What the stand-in cannot pin is nnabla's own kernel arithmetic
(float32 rounding inside cuBLAS / its elementwise kernels) and its RNG streams: random tensors are explicit inputs.
It follows, line by line:
  python/sampler.py:140-165 (stratified), :167-242 (hierarchical), :244-254,:282-291 (background),
  python/network.py:96-117 (positional encoding), :154-232 (geometric), :235-561 (heads, background),
  python/renderer.py:32-209 (pb_render), python/specular_brdf.py:23-118 (filament), python/loss.py:27-192.
nnabla op semantics used (from nnabla's public docs): affine = x @ W + b with W (in,out); softplus(x, beta) =
log(1+exp(beta x))/beta; cumprod(exclusive=True) = shifted product starting at 1; searchsorted(right=False) = first
i with seq[i] >= v; clip_by_value gradient passes inside the range only; nn.grad(outputs, inputs) seeds with ones.
Native kernels on the path (ray bounds, grid query, direction sampling, TV) are taken from oracle/cpu_ref.py, the
grid query re-expressed in torch ops so autograd supplies its first- and second-order derivatives.
"""
import math

import numpy as np
import torch

from . import cpu_ref as R


def T(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ----------------------------------------------------------------------------------------------------
# grid queries in torch ops (same cell arithmetic as cpu_ref.cell; floor is non-differentiable, weights are)
# ----------------------------------------------------------------------------------------------------
def _cell_torch(q, G, r):
    """q (...,3) torch; returns i0,i1 (long), w0,w1 (torch, differentiable wrt q).  The integer cell is computed
    in float32 exactly like the kernels (cpu_ref.cell) so both sides pick the same cell."""
    flat = q.reshape(-1, 3)
    i0, i1, _, _, s, _ = R.cell(flat.detach().to(torch.float32).numpy(), (G, G, G), [-r] * 3, [r] * 3)
    i0 = torch.as_tensor(i0.astype(np.int64)); i1 = torch.as_tensor(i1.astype(np.int64))
    scale = torch.as_tensor(s.astype(np.float64), dtype=q.dtype)
    xyz = (flat - (-r)) * scale
    w0 = i1.to(q.dtype) - xyz            # pqr0 = xyz1 - xyz
    w1 = 1.0 - w0                        # pqr1 = 1 - pqr0
    return i0, i1, w0, w1


def voxel_query_torch(q, F, r=1.0):
    """Trilinear query, reference csrc/grid_feature/voxel_feature_cuda.cu:34-97. F (G,G,G,D)."""
    G = F.shape[0]
    i0, i1, w0, w1 = _cell_torch(q, G, r)
    out = 0
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                ix = (i1 if cx else i0)[:, 0]; iy = (i1 if cy else i0)[:, 1]; iz = (i1 if cz else i0)[:, 2]
                w = (w1 if cx else w0)[:, 0] * (w1 if cy else w0)[:, 1] * (w1 if cz else w0)[:, 2]
                out = out + w[:, None] * F[ix, iy, iz]
    return out.reshape(q.shape[:-1] + (F.shape[-1],))


def triplane_query_torch(q, F, r=1.0):
    """reference triplane_feature_cuda.cu:38-90; F (3,G,G,D) -> (..., D*3), channel c = d*3 + plane."""
    G, D = F.shape[1], F.shape[3]
    i0, i1, w0, w1 = _cell_torch(q, G, r)
    outs = []
    for i, (au, av) in enumerate(((0, 1), (1, 2), (2, 0))):
        f = 0
        for cu in (0, 1):
            for cv in (0, 1):
                u = (i1 if cu else i0)[:, au]; v = (i1 if cv else i0)[:, av]
                w = (w1 if cu else w0)[:, au] * (w1 if cv else w0)[:, av]
                f = f + w[:, None] * F[i][u, v]
        outs.append(f)
    return torch.stack(outs, dim=-1).reshape(q.shape[:-1] + (D * 3,))


def triline_query_torch(q, F, r=1.0):
    """reference triline_feature_cuda.cu:35-83; F (3,G,D) -> (..., D*3)."""
    G, D = F.shape[1], F.shape[2]
    i0, i1, w0, w1 = _cell_torch(q, G, r)
    outs = [w0[:, i][:, None] * F[i][i0[:, i]] + w1[:, i][:, None] * F[i][i1[:, i]] for i in range(3)]
    return torch.stack(outs, dim=-1).reshape(q.shape[:-1] + (D * 3,))


def tv_voxel_torch(q, F, r=1.0):
    """total_variation_loss_cuda.cu:33-85 / :111-174 (backward carries +1e-12 under the rsqrt; here sqrt(x+1e-12)
    differs from it by < 1e-6 relative except for |delta| < 1e-3, documented in the test tolerances)."""
    G = F.shape[0]
    i0, i1, _, _ = _cell_torch(q, G, r)
    f000 = F[i0[:, 0], i0[:, 1], i0[:, 2]]
    dx = F[i1[:, 0], i0[:, 1], i0[:, 2]] - f000
    dy = F[i0[:, 0], i1[:, 1], i0[:, 2]] - f000
    dz = F[i0[:, 0], i0[:, 1], i1[:, 2]] - f000
    return _TVSqrt.apply(dx * dx + dy * dy + dz * dz).reshape(q.shape[:-1] + (F.shape[-1],))


def tv_triplane_torch(q, F, r=1.0):
    G, D = F.shape[1], F.shape[3]
    i0, i1, _, _ = _cell_torch(q, G, r)
    outs = []
    for i, (au, av) in enumerate(((0, 1), (1, 2), (2, 0))):
        f00 = F[i][i0[:, au], i0[:, av]]
        du = F[i][i1[:, au], i0[:, av]] - f00
        dv = F[i][i0[:, au], i1[:, av]] - f00
        outs.append(_TVSqrt.apply(du * du + dv * dv))
    return torch.stack(outs, dim=-1).reshape(q.shape[:-1] + (D * 3,))


def tv_triline_torch(q, F, r=1.0):
    G, D = F.shape[1], F.shape[2]
    i0, i1, _, _ = _cell_torch(q, G, r)
    outs = []
    for i in range(3):
        du = F[i][i1[:, i]] - F[i][i0[:, i]]
        outs.append(_TVSqrt.apply(du * du))
    return torch.stack(outs, dim=-1).reshape(q.shape[:-1] + (D * 3,))


class _TVSqrt(torch.autograd.Function):
    """sqrt(s) forward, 0.5 * rsqrt(s + 1e-12) backward: exactly the kernel pair
    total_variation_loss_cuda.cu:79 (forward) and :161 (backward)."""

    @staticmethod
    def forward(ctx, s):
        ctx.save_for_backward(s)
        return torch.sqrt(s)

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        return g * 0.5 / torch.sqrt(s + 1e-12)


# ----------------------------------------------------------------------------------------------------
# networks (python/network.py)
# ----------------------------------------------------------------------------------------------------
def positional_encoding(x, M):
    """network.py:96-117: [x, cos(b), sin(b)] with b[axis*M + k] = x[axis] * 2^k (axis-major)."""
    if M <= 0:
        return x
    bands = torch.as_tensor(2.0 ** np.arange(M), dtype=x.dtype)
    b = (x[..., None] * bands).reshape(x.shape[:-1] + (-1,))
    return torch.cat([x, torch.cos(b), torch.sin(b)], dim=-1)


def softplus(x, beta):
    # linear for beta*x > 30 (difference from log(1+exp(beta x))/beta < 1e-13/beta) so exp() cannot overflow
    return torch.nn.functional.softplus(x, beta=beta, threshold=30.0)


def mlp(h, layers, act_beta=100.0):
    for (W, b) in layers[:-1]:
        h = softplus(h @ W + b, act_beta)
    W, b = layers[-1]
    return h @ W + b


class Model:
    """Holds the parameters as torch leaves; `P` comes from ndjir_b200.scene.init_params (numpy)."""

    def __init__(self, conf, P, dtype=torch.float64, grids=None):
        self.conf, self.dtype = conf, dtype
        self.nets = {}
        for k, v in P.items():
            if isinstance(v, list):
                self.nets[k] = [(T(W, dtype).requires_grad_(True), T(b, dtype).requires_grad_(True)) for W, b in v]
        self.geo_gain = T(P["geo_gain"], dtype).requires_grad_(True)
        self.pl_gain = T(P["pl_gain"], dtype)
        g = grids if grids is not None else P.get("grid", {})
        self.grid = {k: T(v, dtype).requires_grad_(True) for k, v in g.items() if v is not None}

    def parameters(self):
        out = {}
        off = set()      # networks the configuration switches off are never created by the reference (network.py:308-309,
        if not self.conf.implicit_illumination_network.use_me:                                   # renderer.py:161)
            off.add("ii")
        if not self.conf.photogrammetric_light_network.use_me:
            off.add("pl")
        for k, layers in self.nets.items():
            if k in off:
                continue
            for i, (W, b) in enumerate(layers):
                out[f"{k}.W{i}"], out[f"{k}.b{i}"] = W, b
        out["geo_gain"] = self.geo_gain
        for k, v in self.grid.items():
            out[f"grid.{k}"] = v
        return out

    # network.py:120-151
    def query_on_grid(self, x):
        v = self.conf.geometric_network.voxel
        r = self.conf.renderer.bounding_sphere_radius
        if v.type == "none":
            return None
        if v.use_ste:      # voxel_feature.py:390-391 (and the triplane / triline twins): the backward registered for nn.grad
            x = x.detach()  # returns no gradient for the query - the normal does not see the grid features
        if v.type == "voxel":
            return voxel_query_torch(x, self.grid["voxel"], 1.0)   # PF defaults min=-1,max=1 (voxel_feature.py:147-148)
        if v.type == "triplaneline":
            return torch.cat([triplane_query_torch(x, self.grid["triplane"], 1.0),
                              triline_query_torch(x, self.grid["triline"], 1.0)], dim=-1)
        raise ValueError(v.type)

    # network.py:154-232 (geometric_init branch, softplus beta=100)
    def geometric_network(self, x):
        g = self.conf.geometric_network
        pe_x = positional_encoding(x, g.pe_bands)
        vfeat = self.query_on_grid(x)
        inputs = pe_x if vfeat is None else torch.cat([pe_x, vfeat], dim=-1)
        h = inputs
        L = g.layers
        layers = self.nets["geo"]
        for l in range(L):
            W, b = layers[l]
            h = h @ W + b
            if l == L - 1:
                break
            h = softplus(h, 100.0)
            if g.geometric_init:
                if (l + 1) in g.skip_layers:
                    h = torch.cat([h, inputs], dim=-1)
                    if g.use_inv_square:
                        h = h / math.sqrt(2)
            else:
                raise NotImplementedError("non-geometric-init branch is not on the BASELINE configs")
        sdf, feature = h[..., 0:1], h[..., 1:]
        gain = torch.clamp(torch.exp(self.geo_gain * 10), 1e-6, 5e4)
        return sdf, feature, gain

    def base_color_network(self, x, feature):
        return torch.sigmoid(mlp(torch.cat([x, feature], dim=-1), self.nets["bc"]))          # network.py:235-263

    def environment_light_network(self, dirs):
        c = self.conf.environment_light_network
        h = mlp(positional_encoding(dirs, c.pe_bands), self.nets["el"])                       # network.py:266-297
        return softplus(h, float(c.inverse_black_degree))

    def implicit_illumination_network(self, x, feature, normal):
        return torch.sigmoid(mlp(torch.cat([x, feature, normal], dim=-1), self.nets["ii"]))   # network.py:300-336

    def soft_visibility_light_network(self, x, dirs, feature, normal):
        c = self.conf.soft_visibility_light_network
        h = torch.cat([x, positional_encoding(dirs, c.pe_bands), feature, normal], dim=-1)    # network.py:339-377
        return torch.sigmoid(mlp(h, self.nets["sv"]))

    def photogrammetric_light_network(self, x, camloc, view, feature, normal):
        c = self.conf.photogrammetric_light_network                                          # network.py:380-424
        view = view.expand(x.shape)
        dist2 = ((x - camloc) ** 2).sum(-1, keepdim=True)
        inv = 1.0 / (dist2 + 1e-5)
        parts = [x, positional_encoding(view, c.pe_bands), feature, normal]
        h = torch.cat(parts + ([inv] if c.use_inverse_distance else []), dim=-1)                # network.py:410
        return torch.sigmoid(self.pl_gain * mlp(h, self.nets["pl"]))

    def roughness_network(self, x, feature, normal):
        c = self.conf.roughness_network                                                      # network.py:427-464
        h = mlp(torch.cat([x, feature, normal], dim=-1), self.nets["ro"])
        std = softplus(h[..., 1:2], 1.0)
        r = torch.sigmoid(h[..., 0:1]) ** 2                                                  # filament remap
        return torch.clamp(r, c.lower_bound, 1.0), std

    def specular_reflectance_network(self, x, feature, normal):
        c = self.conf.specular_reflectance_network                                           # network.py:467-509
        h = mlp(torch.cat([x, feature, normal], dim=-1), self.nets["sp"])
        Do = c.channels
        std = softplus(h[..., Do:], 1.0)
        return 0.16 * torch.sigmoid(h[..., :Do]) ** 2, std

    def background_network(self, x, view, delta):
        c = self.conf.background_network                                                     # network.py:512-561
        h = mlp(positional_encoding(x, c.pe_bands0), self.nets["bg0"])
        density, feature = softplus(h[..., 0:1], 100.0), h[..., 1:]
        alpha = 1 - torch.exp(-density * delta)
        view = view.expand(x.shape[:-1] + (3,))
        h = torch.cat([x, feature, view, positional_encoding(view, c.pe_bands1)], dim=-1)
        return alpha, torch.sigmoid(mlp(h, self.nets["bg1"]))


# ----------------------------------------------------------------------------------------------------
# sampler (python/sampler.py)
# ----------------------------------------------------------------------------------------------------
def cumprod_exclusive(x, dim):
    c = torch.cumprod(x, dim=dim)
    one = torch.ones_like(x.narrow(dim, 0, 1))
    return torch.cat([one, c.narrow(dim, 0, x.shape[dim] - 1)], dim=dim)


def importance_round(t, sdf, t_near, t_far, gain, M):
    """One up-sampling round given SDF values at the current samples: sampler.py:196-240.
    t, sdf (B,R,Nt,1); t_near,t_far (B,R,1,1).  Returns the sorted union (B,R,Nt+M,1) and the M new samples."""
    B, Rr, Nt, _ = t.shape
    sdf0, sdf1 = sdf[:, :, :-1], sdf[:, :, 1:]
    t0, t1 = t[:, :, :-1], t[:, :, 1:]
    sdfm = (sdf0 + sdf1) * 0.5
    cos1 = (sdf1 - sdf0) / (t1 - t0 + 1e-5)
    cos0 = torch.cat([torch.ones_like(cos1[:, :, :1]), cos1[:, :, :-1]], dim=2)
    cosv = torch.clamp(torch.minimum(cos0, cos1), -1e3, 0.0)
    dist = t1 - t0
    s0 = sdfm - cosv * dist * 0.5
    s1 = sdfm + cosv * dist * 0.5
    c0, c1 = torch.sigmoid(s0 * gain), torch.sigmoid(s1 * gain)
    alpha = torch.clamp((c0 - c1 + 1e-5) / (c0 + 1e-5), 0.0, 1.0)
    w = (alpha * cumprod_exclusive(1 - alpha, 2))[..., 0]                      # (B,R,Nt-1)
    w = w / w.sum(dim=2, keepdim=True)
    cdf = torch.cumsum(w, dim=2)
    u = (torch.arange(M, dtype=t.dtype) / (M - 1 + 1.0 / M)).reshape(1, 1, M).expand(B, Rr, M).contiguous()
    idx = torch.searchsorted(cdf.contiguous(), u, right=False)                 # first i with cdf[i] >= u
    # q14: idx can equal Nt-1 (out of range for the Nt-1 weights); nnabla's gather on an out-of-range index is
    # undefined, we clamp the gathers on `weights`/`lower` to Nt-2 and keep idx for t/steps
    idx_w = torch.clamp(idx, max=Nt - 2)
    cdf0 = torch.cat([torch.zeros_like(cdf[:, :, :1]), cdf], dim=2)
    denorm = torch.gather(w, 2, idx_w)
    lower = torch.gather(cdf0, 2, idx)
    ratio = ((u - lower) / denorm)[..., None]
    steps = torch.cat([t[:, :, 1:] - t[:, :, :-1], t_far - t[:, :, -1:]], dim=2)
    steps_idx = torch.gather(steps[..., 0], 2, idx)[..., None]
    ts_idx = torch.gather(t[..., 0], 2, idx)[..., None]
    t_new = ts_idx + steps_idx * ratio
    t_new = torch.minimum(torch.maximum(t_new, t_near), t_far)     # clip_by_value = minimum2(maximum2(x, lo), hi)
    t_all, _ = torch.sort(torch.cat([t, t_new], dim=2), dim=2)
    return t_all, t_new, idx


def sample_points(model, camloc, raydir, stratified, background, return_debug=False):
    """SamplePoints._forward_impl, sampler.py:256-299 (no gradient flows through any of it)."""
    conf, dt = model.conf, model.dtype
    r = conf.renderer
    camloc_np, raydir_np = np.asarray(camloc, dtype=np.float32), np.asarray(raydir, dtype=np.float32)
    B, Rr, _ = raydir_np.shape
    rad = r.bounding_sphere_radius
    if r.t_near_far_method == "intersect_with_aabb":
        tn, tf, nh = R.ray_aabb(camloc_np, raydir_np, [-rad] * 3, [rad] * 3)
    elif r.t_near_far_method == "intersect_with_r_sphere":
        tn, tf, nh = R.ray_sphere(camloc_np, raydir_np, rad)
    else:
        raise NotImplementedError(r.t_near_far_method)
    mask = T((nh > 1.0).astype(np.float64), dt).reshape(B, Rr, 1, 1)
    t_near, t_far = T(tn, dt).reshape(B, Rr, 1, 1), T(tf, dt).reshape(B, Rr, 1, 1)
    o = T(camloc_np, dt).reshape(B, 1, 1, 3)
    d = T(raydir_np, dt).reshape(B, Rr, 1, 3)
    N0, M, U = r.n_samples0, r.n_samples1, r.n_upsamples
    with torch.no_grad():
        step = (t_far - t_near) / N0
        t = t_near + step * (torch.arange(N0, dtype=dt).reshape(1, 1, N0, 1) + T(stratified, dt))
        debug = []
        for u in range(U):
            x = o + t * d
            sdf, _, _ = model.geometric_network(x)
            gain = r.sampling_sigmoid_gain * 2 ** u
            t_prev = t
            t, t_new, idx = importance_round(t, sdf, t_near, t_far, gain, M)
            debug.append(dict(t_in=t_prev, sdf=sdf, t_new=t_new, idx=idx, t_out=t))
        x_fg = o + t * d
        t_fg = torch.cat([t, t_far], dim=2)
        if conf.background_modeling:
            cam_d = torch.sqrt((o ** 2).sum(-1, keepdim=True))                 # (B,1,1,1)
            t_near_bg = (cam_d - rad).expand(B, Rr, 1, 1)
            t_base = t_far * mask + t_near_bg * (1 - mask)
            tb, _ = torch.sort(t_base / T(background, dt), dim=2)
            xb = o + tb[:, :, :-1] * d
            dist = torch.sqrt((xb ** 2).sum(-1, keepdim=True)) + 1e-6
            x_bg = torch.cat([xb / dist, 1.0 / dist], dim=-1)
            t_bg = tb
        else:
            Nb = r.n_bg_samples
            x_bg, t_bg = torch.ones(B, Rr, Nb, 4, dtype=dt), torch.ones(B, Rr, Nb + 1, 1, dtype=dt)
    if return_debug:
        return x_fg, t_fg, x_bg, t_bg, mask, debug
    return x_fg, t_fg, x_bg, t_bg, mask


# ----------------------------------------------------------------------------------------------------
# specular BRDF (python/specular_brdf.py:23-118, filament + importance sampling)
# ----------------------------------------------------------------------------------------------------
def dot_clamped(u, v, eps):
    uv = (u * v).sum(-1, keepdim=True)
    mask = (uv > eps).to(u.dtype)
    return torch.clamp(uv, min=eps), mask


def filament_specular_brdf(normal, view_dir, light_dir, roughness, specular_color, conf):
    """normal (B,R,3), view_dir (B,R,1,3), light_dir (B,R,M,3), roughness (B,R,1), specular_color (B,R,C)."""
    n = normal[:, :, None, :].expand(light_dir.shape)
    v = view_dir.expand(light_dir.shape)
    rough = roughness[:, :, None, :]
    sc = specular_color[:, :, None, :]
    half = light_dir + v
    half = half / torch.sqrt((half ** 2).sum(-1, keepdim=True))
    a2 = rough ** 2
    ed = conf.renderer.eps_dot
    nol, m_nol = dot_clamped(n, light_dir, ed)
    nov, m_nov = dot_clamped(n, v, ed)
    noh, m_noh = dot_clamped(n, half, ed)
    eps = 1e-6

    def V1(nou):
        return 1.0 / (nou + torch.sqrt(a2 + (1 - a2) * nou ** 2) + eps)
    voh, _ = dot_clamped(v, half, ed)
    Fs = sc + (1 - sc) * (1 - voh) ** 5
    if conf.specular_brdf.sampling == "importance":
        sBRDF = V1(nol) * V1(nov) * Fs * (4 * voh / noh)
    else:
        D = a2 / (math.pi * (noh ** 2 * (a2 - 1) + 1) ** 2 + eps)
        sBRDF = math.pi * D * V1(nol) * V1(nov) * Fs
    return sBRDF * (m_nol * m_nov * m_noh), nol


# ----------------------------------------------------------------------------------------------------
# renderer (python/renderer.py:32-209) and losses (python/loss.py:27-192)
# ----------------------------------------------------------------------------------------------------
def pb_render(model, x_fg, t_fg, x_bg, t_bg, camloc, raydir, mask, cos_anneal_ratio, rnd, fixed_dirs=None):
    """`fixed_dirs=(dirs_diffuse, dirs_specular)` freezes the (non-differentiable) light directions; used by the
    tests to compare against finite differences and to feed the CUDA path's directions back in."""
    conf, dt = model.conf, model.dtype
    B, Rr, N, _ = x_fg.shape
    raydir = T(raydir, dt).reshape(B, Rr, 1, 3)
    camloc = T(camloc, dt).reshape(B, 1, 1, 3)
    view_dir = -raydir
    sdf, feat, gain = model.geometric_network(x_fg)
    (grad_x,) = torch.autograd.grad(sdf.sum(), x_fg, create_graph=True)        # nn.grad([sdf],[x_fg]) (renderer.py:52)
    c = float(cos_anneal_ratio)
    true_cos = (raydir * grad_x).sum(-1, keepdim=True)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - c) + torch.relu(-true_cos) * c)
    delta = (t_fg[:, :, 1:] - t_fg[:, :, :-1]).detach()
    s1 = sdf + iter_cos * delta * 0.5
    s0 = sdf - iter_cos * delta * 0.5
    c0, c1 = torch.sigmoid(gain * s0), torch.sigmoid(gain * s1)
    alpha_fg = torch.clamp((c0 - c1 + 1e-5) / (c0 + 1e-5), 0.0, 1.0)
    if conf.background_modeling:
        delta_bg = (t_bg[:, :, 1:] - t_bg[:, :, :-1]).detach()
        alpha_bg, color_bg = model.background_network(x_bg, view_dir, delta_bg)
    else:
        alpha_bg = torch.ones(B, Rr, 1, 1, dtype=dt)
        color_bg = torch.full((B, Rr, 1, 3), float(conf.background_color), dtype=dt)
    alpha = torch.cat([alpha_fg * mask, alpha_bg], dim=2)
    trans = cumprod_exclusive(1 - alpha, 2)
    weights = alpha * trans
    w_fg, w_bg = weights[:, :, :N], weights[:, :, N:]
    trans_fg = trans[:, :, :N]

    def VR(x, w=w_fg):
        return (w * x).sum(dim=2)
    g_pix = VR(grad_x) + conf.renderer.eps_normal
    normal_pixel = g_pix / torch.sqrt((g_pix ** 2).sum(-1, keepdim=True))
    nt = conf.renderer.n_thetas
    M = nt * 2 * nt
    x_pix = VR(x_fg)[:, :, None, :].expand(B, Rr, M, 3)
    f_pix = VR(feat)[:, :, None, :].expand(B, Rr, M, feat.shape[-1])
    n_b = normal_pixel[:, :, None, :].expand(B, Rr, M, 3)
    # diffuse directions (no gradient through the sampler: SampleDirections.backward_impl is empty, sampler.py:391)
    if fixed_dirs is not None:
        dirs_u = T(fixed_dirs[0], dt)
    else:
        dirs_u = T(R.sample_directions(normal_pixel.detach().to(torch.float32).numpy(), rnd["diffuse_cdf_the"],
                                       rnd["diffuse_cdf_phi"]), dt)
    env = model.environment_light_network(dirs_u)
    vis = model.soft_visibility_light_network(x_pix, dirs_u, f_pix, n_b)
    if conf.implicit_illumination_network.use_me:
        ii = model.implicit_illumination_network(x_fg, feat, grad_x)
    else:                                                   # network.py:308-309: a constant 0, no parameters
        ii = torch.zeros(x_fg.shape[:-1] + (1,), dtype=dt)
    ii_pix = VR(ii)
    cosd, _ = dot_clamped(n_b, dirs_u, 1e-8)                                   # specular_brdf.dot default eps
    env_diffuse = (vis * env * cosd).mean(dim=2)
    diffuse_light = env_diffuse + ii_pix
    base_color = model.base_color_network(x_fg, feat)
    rough, std_r = model.roughness_network(x_fg, feat, grad_x)
    rough_pix = VR(rough)
    spec, std_s = model.specular_reflectance_network(x_fg, feat, grad_x)
    spec_pix = VR(spec)
    if fixed_dirs is not None:
        dirs_s = T(fixed_dirs[1], dt)
    elif conf.specular_brdf.sampling == "importance":
        dirs_s = T(R.sample_directions(normal_pixel.detach().to(torch.float32).numpy(), rnd["specular_cdf_the"],
                                       rnd["specular_cdf_phi"], rough_pix.detach().to(torch.float32).numpy()), dt)
    else:
        dirs_s = T(R.sample_directions(normal_pixel.detach().to(torch.float32).numpy(), rnd["specular_cdf_the"],
                                       rnd["specular_cdf_phi"]), dt)
    sBRDF, cos_s = filament_specular_brdf(normal_pixel, view_dir, dirs_s, rough_pix, spec_pix, conf)
    env_s = model.environment_light_network(dirs_s)
    vis_s = model.soft_visibility_light_network(x_pix, dirs_s, f_pix, n_b)
    spec_color = (sBRDF * vis_s * env_s * cos_s).mean(dim=2) * conf.specular_brdf.weight
    if conf.photogrammetric_light_network.use_me:
        pl = model.photogrammetric_light_network(x_fg, camloc, view_dir, feat, grad_x)
        pl_pix = VR(pl)
        if conf.diffuse_brdf.entangle:
            color_fg = VR(base_color * pl) * diffuse_light + pl_pix * spec_color
        else:
            color_fg = pl_pix * (VR(base_color) * diffuse_light + spec_color)
    else:                                                   # renderer.py:174-176 (sic: no diffuse-light factor)
        pl = torch.ones(x_fg.shape[:-1] + (1,), dtype=dt)
        color_fg = VR(base_color) + spec_color
    color = color_fg + VR(color_bg, w_bg)
    # colour perturbation branch (renderer.py:187-193)
    G = conf.geometric_network.voxel.grid_size
    rad = conf.renderer.bounding_sphere_radius
    x_ptb = x_fg.detach() + T(rnd["perturb"], dt) * (math.sqrt(3) * 2 * rad / G)
    _, feat_ptb, _ = model.geometric_network(x_ptb)
    bc_ptb = model.base_color_network(x_ptb, feat_ptb)
    return dict(color_pixel=color, sdf_x_fg=sdf, grad_x_fg=grad_x, alpha_fg=alpha_fg, trans_fg=trans_fg,
                base_color=base_color, base_color_ptb=bc_ptb, roughness=rough, specular_reflectance=spec,
                std_roughness=std_r, std_specular_reflectance=std_s, weights_fg=w_fg, weights_bg=w_bg,
                normal_pixel=normal_pixel, feature=feat, implicit=ii, photogrammetric=pl, color_bg=color_bg,
                alpha_bg=alpha_bg, dirs_diffuse=dirs_u, dirs_specular=dirs_s, diffuse_light=diffuse_light,
                specular_color=spec_color, roughness_pixel=rough_pix, specular_pixel=spec_pix)


def total_loss(model, camloc, raydir, color_gt, cos_anneal_ratio, rnd, return_all=False, samples=None,
               fixed_dirs=None, obj_mask=None):
    """loss.py:27-192.  `samples` = (x_fg,t_fg,x_bg,t_bg,mask) overrides sample_points (tests freeze the
    non-differentiable placement).  obj_mask (B,R,1) is read when train.mask_weight > 0 (loss.py:108-116)."""
    conf, dt = model.conf, model.dtype
    tr = conf.train
    if samples is None:
        samples = sample_points(model, camloc, raydir, rnd["stratified"], rnd["background"])
    x_fg, t_fg, x_bg, t_bg, mask = [T(v, dt) if not torch.is_tensor(v) else v.to(dt) for v in samples]
    x_fg = x_fg.detach().requires_grad_(True)
    res = pb_render(model, x_fg, t_fg, x_bg, t_bg, camloc, raydir, mask, cos_anneal_ratio, rnd, fixed_dirs)
    B, Rr, N, _ = x_fg.shape
    gt = T(color_gt, dt)
    per_ray = (res["color_pixel"] - gt).abs() if tr.rgb_loss == "l1" else (res["color_pixel"] - gt) ** 2
    if tr.mask_weight > 0.0:      # loss.py:63-65: with a mask term the colour loss covers the object's rays only
        om = T(obj_mask, dt).reshape(B, Rr, 1)
        loss_rgb = (per_ray * om).sum() / (om.sum() + 1e-5)
    else:
        loss_rgb = per_ray.sum() / (B * Rr)
    denorm = mask.sum() * N + 1e-5
    losses = {"loss_rgb": loss_rgb}
    gn = torch.sqrt((res["grad_x_fg"] ** 2).sum(-1, keepdim=True))
    losses["loss_eikonal"] = (((gn - 1) * mask) ** 2).sum() / denorm
    zero = torch.zeros((), dtype=dt)
    loss_tv = zero
    if conf.geometric_network.voxel.type != "none" and tr.tv_weight > 0:
        for name, F in model.grid.items():
            tvf = {"voxel": tv_voxel_torch, "triplane": tv_triplane_torch, "triline": tv_triline_torch}[name]
            tv = _tv_with_sym(tvf, x_fg.detach(), F, tr.tv_sym_backward)
            loss_tv = loss_tv + (tv * mask).sum() / denorm
    losses["loss_tv"] = loss_tv
    bc = res["base_color"] if tr.base_color_prior_sym_backward else res["base_color"].detach()
    losses["prior_base_color"] = ((bc - res["base_color_ptb"]).abs() * mask).sum() / denorm
    ro, sr = res["roughness"], res["std_roughness"]
    losses["prior_roughness"] = (((ro - conf.roughness_network.prior_value).abs() / sr) * mask).sum() / denorm
    losses["reg_std_roughness"] = (torch.clamp(torch.log(sr), 1e-5, 1e5) * mask).sum() / denorm
    sp, ss = res["specular_reflectance"], res["std_specular_reflectance"]
    losses["prior_specular_reflectance"] = (((sp - conf.specular_reflectance_network.prior_value).abs() / ss)
                                            * mask).sum() / denorm
    losses["reg_std_specular_reflectance"] = (torch.clamp(torch.log(ss), 1e-5, 1e5) * mask).sum() / denorm
    # mask loss (loss.py:108-116): BCE of the clipped opacity sum_i alpha_i T_i (renderer.py:183-185) against the mask
    losses["loss_mask"] = zero
    if tr.mask_weight > 0.0:
        # renderer.py:183-185: the UNMASKED alpha_fg times the transmittance of the masked alphas
        pred = torch.clamp((res["alpha_fg"] * res["trans_fg"]).sum(dim=2), 1e-3, 1.0 - 1e-3)
        y = T(obj_mask, dt).reshape(pred.shape)
        losses["loss_mask"] = (-(y * torch.log(pred) + (1 - y) * torch.log(1 - pred))).sum() / (mask.sum() + 1e-5)
    loss = (losses["loss_rgb"] + tr.eikonal_weight * losses["loss_eikonal"] + tr.tv_weight * losses["loss_tv"]
            + tr.mask_weight * losses["loss_mask"]
            + tr.base_color_prior_weight * losses["prior_base_color"]
            + tr.roughness_prior_weight * (losses["prior_roughness"] + losses["reg_std_roughness"])
            + tr.specular_reflectance_prior_weight * (losses["prior_specular_reflectance"]
                                                      + losses["reg_std_specular_reflectance"]))
    losses["loss"] = loss
    # a term whose weight is not > 0 is never built by the reference and reads 0.0 in its dict (loss.py:70-165)
    for w, keys in ((tr.eikonal_weight, ("loss_eikonal",)), (tr.base_color_prior_weight, ("prior_base_color",)),
                    (tr.roughness_prior_weight, ("prior_roughness", "reg_std_roughness")),
                    (tr.specular_reflectance_prior_weight, ("prior_specular_reflectance", "reg_std_specular_reflectance"))):
        if not w > 0.0:
            for k in keys:
                losses[k] = zero
    if return_all:
        return losses, res, dict(x_fg=x_fg, t_fg=t_fg, x_bg=x_bg, t_bg=t_bg, mask=mask)
    return losses


def _tv_with_sym(tvf, x, F, sym):
    """sym_backward=False drops the gradient into the corner cell itself (total_variation_loss_cuda.cu:169-172):
    emulated by evaluating the corner term on a detached copy of the table."""
    if sym:
        return tvf(x, F)
    # gradient only through the neighbour cells: f(F) with the f000 path detached
    return _TVNoSym.apply(x, F, tvf)


class _TVNoSym(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, F, tvf):
        ctx.tvf, ctx.x = tvf, x
        ctx.save_for_backward(F)
        return tvf(x, F.detach())

    @staticmethod
    def backward(ctx, g):
        (F,) = ctx.saved_tensors
        fam = {tv_voxel_torch: R.tv_voxel_backward, tv_triplane_torch: R.tv_triplane_backward,
               tv_triline_torch: R.tv_triline_backward}[ctx.tvf]
        gf = fam(g.detach().numpy().reshape(-1, g.shape[-1]), ctx.x.detach().numpy().reshape(-1, 3),
                 F.detach().to(torch.float32).numpy(), [-1.0] * 3, [1.0] * 3, False)
        return None, torch.as_tensor(gf, dtype=F.dtype), None


def train_step(model, camloc, raydir, color_gt, cos_anneal_ratio, rnd, samples=None, fixed_dirs=None, obj_mask=None):
    """loss.forward() + loss.backward() (python/train.py:135-140): returns ({loss terms}, {param: grad})."""
    params = model.parameters()
    for p in params.values():
        p.grad = None
    losses = total_loss(model, camloc, raydir, color_gt, cos_anneal_ratio, rnd, samples=samples, fixed_dirs=fixed_dirs,
                        obj_mask=obj_mask)
    losses["loss"].backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    return {k: float(v.detach()) for k, v in losses.items()}, grads
