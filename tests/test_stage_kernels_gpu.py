"""GPU: the stage kernels of the fused per-ray path, called DIRECTLY through the C ABI at the per-ray shapes of
default.yaml (64/80/96/112 -> +16 samples, 128 + 32 composited samples = 5 scan chunks, 33 sorted background
distances, 128 light directions), each against a restatement of the reference lines it replaces.

  ndjir_importance_round{,_incremental}   python/sampler.py:196-240   float32 numpy restatement in the kernel's op order
  ndjir_background_samples                python/sampler.py:244-254, 282-291
  ndjir_composite_{forward,backward}      python/renderer.py:79-84 (+ autograd of it in float64)
  ndjir_shade_{forward,backward}          python/renderer.py:93-180, specular_brdf.py:40-118 (oracle/cpu_render.py, pinned on
                                          the reference's own Python by tests/test_render_golden.py)
Bars: sample indices bit-exact wherever the decision is not inside float32 rounding of the CDF (and >= 99.9 % overall),
sorted unions bit-exact given the new distances, distances 1e-6 (+ the first-order bound of float32 CDF rounding, 2e-6, through
(u - cdf)/w for sections with w ~ 1e-5), forward values 1e-5, backward 1e-4."""
import numpy as np
import pytest
import torch

from ndjir_b200 import _lib
from ndjir_b200.config import make_conf
from oracle import cpu_ref as R
from oracle import cpu_render as CR

pytestmark = pytest.mark.gpu
f32 = np.float32


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).cuda()


def call(name, *args):
    _lib.call(name, *args, torch.cuda.current_stream().cuda_stream)


def host(t):
    return t.detach().cpu().numpy()


# ----------------------------------------------------------------------------------------------------
# float32 restatement of one up-sampling round, op for op as csrc/sampler.cu evaluates it (scans included)
# ----------------------------------------------------------------------------------------------------
def _warp_scan(x, op):
    """Hillis-Steele inclusive scan over the last axis (32 lanes), the kernel's shuffle order."""
    x = x.copy()
    for o in (1, 2, 4, 8, 16):
        up = x[..., :-o].copy()
        x[..., o:] = op(x[..., o:], up)
    return x


def _warp_sum(x):
    idx = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        x = x + x[..., idx ^ o]
    return x[..., 0]


def importance_round_f32(t, sdf, tn, tf, gain, M):
    """t, sdf (NR, Nt) float32 sorted by t; returns idx (NR, M) int, t_new (NR, M) f32, cdf, w (for conditioning)."""
    NR, Nt = t.shape
    S = Nt - 1
    one, half, e5 = f32(1), f32(0.5), f32(1e-5)
    gain = f32(gain)
    s0v, s1v, t0, t1 = sdf[:, :-1], sdf[:, 1:], t[:, :-1], t[:, 1:]
    mid = (s0v + s1v) * half
    cos1 = (s1v - s0v) / (t1 - t0 + e5)
    cos0 = np.concatenate([np.ones((NR, 1), f32), cos1[:, :-1]], axis=1)
    c = np.minimum(cos0, cos1)
    c = np.minimum(np.maximum(c, f32(-1e3)), f32(0))
    dist = t1 - t0
    a0 = mid - c * dist * half
    a1 = mid + c * dist * half
    with np.errstate(over="ignore"):
        c0 = one / (one + np.exp(-(a0 * gain), dtype=f32))
        c1 = one / (one + np.exp(-(a1 * gain), dtype=f32))
    alpha = (c0 - c1 + e5) / (c0 + e5)
    alpha = np.minimum(np.maximum(alpha, f32(0)), one)
    nch = (S + 31) // 32
    ap = np.zeros((NR, nch * 32), f32)
    ap[:, :S] = alpha
    w = np.zeros_like(ap)
    carry = np.ones(NR, f32)
    total_l = np.zeros((NR, 32), f32)
    for ch in range(nch):
        a = ap[:, ch * 32:(ch + 1) * 32]
        incl = _warp_scan(one - a, np.multiply)
        excl = np.concatenate([np.ones((NR, 1), f32), incl[:, :-1]], axis=1)
        wc = a * carry[:, None] * excl
        valid = (np.arange(32) + ch * 32) < S
        wc = np.where(valid, wc, f32(0))
        w[:, ch * 32:(ch + 1) * 32] = wc
        total_l = total_l + wc
        carry = carry * incl[:, 31]
    total = _warp_sum(total_l)
    cdf = np.zeros_like(ap)
    run = np.zeros(NR, f32)
    for ch in range(nch):
        valid = (np.arange(32) + ch * 32) < S
        wn = np.where(valid, w[:, ch * 32:(ch + 1) * 32] / total[:, None], f32(0))
        incl = _warp_scan(wn, np.add)
        w[:, ch * 32:(ch + 1) * 32] = wn
        cdf[:, ch * 32:(ch + 1) * 32] = run[:, None] + incl
        run = run + incl[:, 31]
    w, cdf = w[:, :S], cdf[:, :S]
    k = np.arange(M, dtype=f32)
    u = k / (f32(M - 1) + one / f32(M))
    idx = np.stack([np.searchsorted(cdf[r], u, side="left") for r in range(NR)])          # first i with cdf[i] >= u
    idx_w = np.minimum(idx, S - 1)
    rows = np.arange(NR)[:, None]
    lower = np.where(idx == 0, f32(0), cdf[rows, np.maximum(idx - 1, 0)])
    ratio = (u[None, :] - lower) / w[rows, idx_w]
    steps = np.concatenate([t[:, 1:] - t[:, :-1], tf[:, None] - t[:, -1:]], axis=1)
    t_new = t[rows, np.minimum(idx, Nt - 1)] + steps[rows, np.minimum(idx, Nt - 1)] * ratio
    t_new = np.minimum(np.maximum(t_new, tn[:, None]), tf[:, None])
    return idx.astype(np.int32), t_new.astype(f32), cdf, w, u, steps


def _rays(NR, seed):
    rng = np.random.RandomState(seed)
    conf = make_conf("default")
    o = rng.randn(NR, 3)
    o = (o / np.linalg.norm(o, axis=1, keepdims=True) * rng.uniform(2.5, 3.0, (NR, 1))).astype(f32)
    target = rng.uniform(-0.6, 0.6, (NR, 3))
    d = target - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    tn, tf, nh = R.ray_aabb(o[:, None, :].reshape(NR, 3), d.reshape(NR, 1, 3), [-1.0] * 3, [1.0] * 3)
    return conf, rng, o, d, np.asarray(tn, f32).reshape(NR), np.asarray(tf, f32).reshape(NR)


def _scene_sdf(o, d, t, rng):
    """a sphere of radius 0.5 plus ripples: the SDF profile along each ray (float32, as the network would hand it over)"""
    x = o[:, None, :] + t[:, :, None] * d[:, None, :]
    s = np.linalg.norm(x, axis=-1) - 0.5 + 0.03 * np.sin(9.0 * x[..., 0]) * np.cos(7.0 * x[..., 1])
    return s.astype(f32)


@pytest.mark.parametrize("u_round", [0, 1, 2, 3])
def test_importance_round_matches_fp32_restatement(u_round):
    NR, N0, M = 512, 64, 16
    Nt = N0 + u_round * M
    conf, rng, o, d, tn, tf = _rays(NR, 100 + u_round)
    # current samples: stratified ones plus u_round*M extra ones concentrated near the surface, sorted
    xi = rng.rand(NR, N0).astype(f32)
    t = tn[:, None] + (tf - tn)[:, None] / f32(N0) * (np.arange(N0, dtype=f32)[None] + xi)
    if u_round:
        extra = (tn[:, None] + (tf - tn)[:, None] * rng.beta(8, 8, (NR, u_round * M))).astype(f32)
        t = np.concatenate([t, extra], axis=1)
    t = np.sort(t.astype(f32), axis=1)
    sdf = _scene_sdf(o, d, t, rng)
    gain = conf.renderer.sampling_sigmoid_gain * 2 ** u_round
    t_out = torch.empty(NR, Nt + M, device="cuda")
    t_new = torch.empty(NR, M, device="cuda")
    idx = torch.empty(NR, M, dtype=torch.int32, device="cuda")
    call("ndjir_importance_round", NR, Nt, M, dev(t), Nt, dev(sdf), Nt, dev(tn), dev(tf), float(gain), t_out, Nt + M,
         t_new, idx)
    torch.cuda.synchronize()
    g_idx, g_new, g_out = host(idx), host(t_new), host(t_out)
    r_idx, r_new, cdf, w, uq, steps = importance_round_f32(t, sdf, tn, tf, gain, M)
    # indices: exact wherever u is not within float32 rounding of a CDF entry next to the chosen one
    rows = np.arange(NR)[:, None]
    S = Nt - 1
    below = np.where(r_idx > 0, cdf[rows, np.maximum(r_idx - 1, 0)], -1.0)
    at = cdf[rows, np.minimum(r_idx, S - 1)]
    margin = np.minimum(np.abs(uq[None] - below), np.where(r_idx < S, np.abs(at - uq[None]), 1.0))
    decided = margin > 2e-5
    assert np.array_equal(g_idx[decided], r_idx[decided]), "sample index differs where the CDF decision is clear"
    assert (g_idx == r_idx).mean() >= 0.999, (g_idx == r_idx).mean()
    assert decided.mean() > 0.9
    # new distances: 1e-6 relative where well conditioned; one ulp of CDF error moves t by step/w * 6e-8
    same = g_idx == r_idx
    wsel = w[rows, np.minimum(r_idx, S - 1)]
    bound = 1e-6 * np.abs(r_new).max() + 2e-6 * np.abs(steps[rows, np.minimum(r_idx, Nt - 1)]) / np.maximum(wsel, 1e-30)
    err = np.abs(g_new - r_new)
    assert (err[same] <= bound[same]).all(), float((err[same] / bound[same]).max())
    well = same & (wsel > 1e-2)
    assert well.mean() > 0.2 and (err[well] <= 3e-6 * np.abs(r_new).max()).all()
    # the sorted union is exactly sort(concat(t, t_new)) of the kernel's own new distances
    assert np.array_equal(g_out, np.sort(np.concatenate([t, g_new], axis=1), axis=1))
    assert (g_new >= tn[:, None]).all() and (g_new <= tf[:, None]).all()


def test_importance_round_incremental_equals_full_round():
    """Rounds chained the engine's way (carry (t, sdf) pairs, merge the pending samples with the key-value sort) give
    bit-identical placements to re-running the full round on the merged arrays."""
    NR, N0, M, U = 256, 64, 16, 4
    conf, rng, o, d, tn, tf = _rays(NR, 7)
    xi = rng.rand(NR, N0).astype(f32)
    pend = (tn[:, None] + (tf - tn)[:, None] / f32(N0) * (np.arange(N0, dtype=f32)[None] + xi)).astype(f32)
    N = N0 + U * M
    t_cur = torch.zeros(NR, N + 1, device="cuda")
    s_cur = torch.zeros(NR, N, device="cuda")
    Nt, Mp = 0, N0
    t_ref = np.zeros((NR, 0), f32)
    for u in range(U + 1):
        last = u == U
        sdf_p = _scene_sdf(o, d, pend, rng)
        gain = float(conf.renderer.sampling_sigmoid_gain * 2 ** u)
        tnew = torch.empty(NR, M, device="cuda")
        idx = torch.empty(NR, M, dtype=torch.int32, device="cuda")
        call("ndjir_importance_round_incremental", NR, Nt, Mp, 0 if last else M, t_cur, N + 1, s_cur, N, dev(pend),
             dev(sdf_p), dev(tn), dev(tf), gain, tnew, idx)
        torch.cuda.synchronize()
        t_ref = np.sort(np.concatenate([t_ref, pend], axis=1), axis=1)
        Nt += Mp
        merged_t, merged_s = host(t_cur)[:, :Nt], host(s_cur)[:, :Nt]
        assert np.array_equal(merged_t, t_ref), f"round {u}: merged distances"
        assert np.array_equal(merged_s, _scene_sdf(o, d, merged_t, rng)), f"round {u}: sdf values follow their samples"
        if last:
            break
        t_out = torch.empty(NR, Nt + M, device="cuda")
        tnew2 = torch.empty(NR, M, device="cuda")
        idx2 = torch.empty(NR, M, dtype=torch.int32, device="cuda")
        call("ndjir_importance_round", NR, Nt, M, dev(merged_t), Nt, dev(merged_s), Nt, dev(tn), dev(tf), gain, t_out,
             Nt + M, tnew2, idx2)
        torch.cuda.synchronize()
        assert torch.equal(idx, idx2) and torch.equal(tnew, tnew2), f"round {u}"
        pend, Mp = host(tnew), M
    assert Nt == N


def test_background_samples_33_sorted():
    NR, Nb, Rr = 300, 32, 75
    conf, rng, o4, d, tn, tf = _rays(NR, 11)
    cam = o4[::Rr][:NR // Rr].copy()                     # (B,3): rays of a view share the camera
    o = np.repeat(cam, Rr, axis=0)
    mask = (rng.rand(NR) > 0.2).astype(f32)
    xi = (rng.rand(NR, Nb + 1).astype(f32) * f32(1 - 1e-5) + f32(1e-5))
    t_bg = torch.empty(NR, Nb + 1, device="cuda")
    x_bg = torch.empty(NR, Nb, 4, device="cuda")
    call("ndjir_background_samples", NR, Nb, Rr, dev(cam), dev(d), dev(tf), dev(mask), dev(xi), 1.0, t_bg, x_bg)
    torch.cuda.synchronize()
    t_base64 = tf.astype(np.float64) * mask + (np.linalg.norm(o.astype(np.float64), axis=1) - 1.0) * (1 - mask)
    want_t = np.sort(t_base64[:, None] / xi, axis=1)
    got_t = host(t_bg)
    assert np.all(np.diff(got_t, axis=1) >= 0)
    np.testing.assert_allclose(got_t, want_t, rtol=6e-7, atol=0)
    p = o[:, None, :].astype(np.float64) + got_t[:, :Nb, None].astype(np.float64) * d[:, None, :]
    dist = np.linalg.norm(p, axis=-1, keepdims=True) + 1e-6
    want_x = np.concatenate([p / dist, 1.0 / dist], axis=-1)
    np.testing.assert_allclose(host(x_bg), want_x, rtol=2e-6, atol=2e-7)


def test_composite_forward_backward_160_samples():
    NR, N, Nb = 97, 128, 32
    S = N + Nb
    rng = np.random.RandomState(3)
    a_fg = (rng.rand(NR, N) ** 4 * 0.6).astype(f32)
    a_fg[rng.rand(NR, N) < 0.3] = 0.0
    a_fg[5, 40] = 1.0                                     # an opaque sample: everything behind it has T = 0
    a_bg = (rng.rand(NR, Nb) * 0.5).astype(f32)
    mask = (rng.rand(NR) > 0.15).astype(f32)
    dw = rng.randn(NR, S).astype(f32)
    w, T = torch.empty(NR, S, device="cuda"), torch.empty(NR, S, device="cuda")
    call("ndjir_composite_forward", NR, N, Nb, dev(a_fg), dev(mask), dev(a_bg), w, T)
    da_fg, da_bg = torch.empty(NR, N, device="cuda"), torch.empty(NR, Nb, device="cuda")
    call("ndjir_composite_backward", NR, N, Nb, dev(a_fg), dev(mask), dev(a_bg), T, dev(dw), da_fg, da_bg)
    torch.cuda.synchronize()
    afg = torch.tensor(a_fg, dtype=torch.float64, requires_grad=True)
    abg = torch.tensor(a_bg, dtype=torch.float64, requires_grad=True)
    alpha = torch.cat([afg * torch.tensor(mask, dtype=torch.float64)[:, None], abg], dim=1)
    Tr = CR.cumprod_exclusive(1 - alpha, 1)                # renderer.py:80
    wr = alpha * Tr
    (wr * torch.tensor(dw, dtype=torch.float64)).sum().backward()
    np.testing.assert_allclose(host(T), Tr.detach().numpy(), rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(host(w), wr.detach().numpy(), rtol=2e-6, atol=1e-9)
    assert np.allclose((host(w).sum(1) + host(T)[:, -1] * (1 - a_bg[:, -1])), 1.0, atol=1e-5)   # mass conservation
    for got, want in ((da_fg, afg.grad), (da_bg, abg.grad)):
        want = want.numpy()
        assert np.abs(host(got) - want).max() <= 1e-5 * np.abs(want).max()


def test_shade_forward_backward_128_directions():
    NR, M, nt = 40, 128, 8
    rng = np.random.RandomState(9)
    conf = make_conf("default")
    nhat = rng.randn(NR, 3)
    nhat = (nhat / np.linalg.norm(nhat, axis=1, keepdims=True)).astype(f32)
    view = nhat + 0.8 * rng.randn(NR, 3)                  # mostly in front of the surface, a few grazing / behind
    view = view / np.linalg.norm(view, axis=1, keepdims=True)
    raydir = (-view).astype(f32)
    att = np.zeros((NR, 12), f32)
    att[:, 0] = rng.rand(NR)
    att[:, 1] = rng.uniform(0.089, 1.0, NR)
    att[:, 2:5] = rng.uniform(0.0, 0.16, (NR, 3))
    att[:, 5] = rng.rand(NR)
    att[:, 6:9] = rng.rand(NR, 3)
    dirs_u = R.sample_directions(nhat.reshape(1, NR, 3), rng.rand(1, NR, nt).astype(f32),
                                 rng.rand(1, NR, 2 * nt).astype(f32)).reshape(NR, M, 3).astype(f32)
    dirs_s = R.sample_directions(nhat.reshape(1, NR, 3), rng.rand(1, NR, nt).astype(f32),
                                 rng.rand(1, NR, 2 * nt).astype(f32), att[:, 1].reshape(1, NR, 1)).reshape(NR, M, 3).astype(f32)
    el_raw = np.zeros((2 * NR * M, 4), f32)
    sv_raw = np.zeros((2 * NR * M, 4), f32)
    el_raw[:, 0] = rng.randn(2 * NR * M)
    sv_raw[:, 0] = rng.randn(2 * NR * M)
    colbg = (rng.rand(NR, 3) * 0.2).astype(f32)
    gt = rng.rand(NR, 3).astype(f32)
    inv_rays = 1.0 / NR
    cfg5 = [conf.renderer.eps_dot, conf.specular_brdf.weight, inv_rays, 1.0, 0.0]
    color = torch.empty(NR, 3, device="cuda")
    losses = torch.zeros(16, device="cuda")
    args = (NR, M, dev(nhat), dev(att), dev(raydir), dev(dirs_u), dev(dirs_s), dev(el_raw), 4, dev(sv_raw), 4,
            dev(colbg), dev(gt), cfg5)
    call("ndjir_shade_forward", *args, color, losses)
    d_el, d_sv = torch.zeros(2 * NR * M, 4, device="cuda"), torch.zeros(2 * NR * M, 4, device="cuda")
    d_att, d_nh, d_cb = torch.empty(NR, 12, device="cuda"), torch.empty(NR, 3, device="cuda"), torch.empty(NR, 3, device="cuda")
    call("ndjir_shade_backward", *args, d_el, d_sv, d_att, d_nh, d_cb)
    torch.cuda.synchronize()
    # float64 restatement of renderer.py:93-180 with oracle/cpu_render's filament BRDF (pinned by test_render_golden)
    D = torch.float64

    def leaf(a):
        return torch.tensor(np.asarray(a, np.float64), dtype=D, requires_grad=True)
    n, A, el, sv, cb = leaf(nhat), leaf(att), leaf(el_raw[:, 0]), leaf(sv_raw[:, 0]), leaf(colbg)
    nB = n.reshape(1, NR, 3)
    v = -torch.tensor(raydir.astype(np.float64)).reshape(1, NR, 1, 3)
    du, ds = torch.tensor(dirs_u.astype(np.float64)).reshape(1, NR, M, 3), torch.tensor(dirs_s.astype(np.float64)).reshape(1, NR, M, 3)
    env = CR.softplus(el.reshape(2, 1, NR, M, 1), 1.0)
    vis = torch.sigmoid(sv.reshape(2, 1, NR, M, 1))
    n_b = nB[:, :, None, :].expand(1, NR, M, 3)
    cosd, _ = CR.dot_clamped(n_b, du, 1e-8)
    Ed = (vis[0] * env[0] * cosd).mean(dim=2)
    DL = Ed + A[:, 0].reshape(1, NR, 1)
    sBRDF, nol = CR.filament_specular_brdf(nB, v, ds, A[:, 1].reshape(1, NR, 1), A[:, 2:5].reshape(1, NR, 3), conf)
    Sp = (sBRDF * vis[1] * env[1] * nol).mean(dim=2) * conf.specular_brdf.weight
    col = A[:, 6:9].reshape(1, NR, 3) * DL + A[:, 5].reshape(1, NR, 1) * Sp + cb.reshape(1, NR, 3)
    loss = (col - torch.tensor(gt.astype(np.float64)).reshape(1, NR, 3)).abs().sum()
    (loss * inv_rays).backward()
    want_col = col.detach().numpy().reshape(NR, 3)
    assert np.abs(host(color) - want_col).max() <= 1e-5 * np.abs(want_col).max()
    assert abs(float(losses[1]) - float(loss)) <= 1e-5 * float(loss)

    def close(got, want, what, tol=1e-4):
        want = np.asarray(want, np.float64)
        e = np.abs(host(got).astype(np.float64).reshape(want.shape) - want).max() / max(np.abs(want).max(), 1e-30)
        assert e <= tol, (what, e)
    close(d_el[:, 0], el.grad.numpy(), "d el_raw")
    close(d_sv[:, 0], sv.grad.numpy(), "d sv_raw")
    close(d_att[:, :9], A.grad.numpy()[:, :9], "d attpix")
    close(d_nh, n.grad.numpy(), "d nhat")
    close(d_cb, cb.grad.numpy(), "d colbg")


@pytest.mark.parametrize("N,Nb,C,cos_anneal", [(128, 32, 262, 0.0), (128, 32, 262, 0.7), (16, 4, 70, 1.0), (40, 0, 9, 0.3)])
def test_render_segment_fused_equals_stage_kernels(N, Nb, C, cos_anneal):
    """ndjir_render_segment_{forward,backward} (ONE kernel per ray: alpha, background alpha, scan, weights, reductions,
    background colour, and the whole backward of that stage) against the separate stage kernels it fuses
    (ndjir_neus_alpha_*, ndjir_bg_alpha_*, ndjir_composite_*, ndjir_volume_render_*, ndjir_bg_color_*), which the tests
    above and the engine parity tests check against the reference's arithmetic: forward bit-identical, backward 1e-6."""
    rng = np.random.RandomState(11)
    NR, C2 = 37, 12
    P, S, rb = NR * N, N + Nb, NR * max(Nb, 1)
    ld_v = C + 2
    sdf = dev(rng.randn(P, 1) * 0.05)
    nrm = dev(rng.randn(P, 3))
    rd = rng.randn(NR, 3); rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    raydir = dev(rd)
    t_fg = dev(np.sort(rng.rand(NR, N + 1) * 2 + 1, axis=1))
    t_bg = dev(np.sort(rng.rand(NR, Nb + 1) * 20 + 3, axis=1))
    gain = dev(np.array([0.3, 0, 0, 0]))
    mask = dev((rng.rand(NR) > 0.2).astype(f32))
    h0 = dev(rng.randn(rb, 1) * 0.02)
    raw = dev(rng.randn(rb, 4))
    V = dev(rng.randn(P, ld_v))
    V2 = dev(rng.randn(P, C2))
    z = lambda *s: torch.zeros(s, device="cuda")
    # ---- separate kernels ----
    a_fg, a_bg, w, T, pix, colbg = z(P, 1), z(rb, 1), z(NR, S), z(NR, S), z(NR, ld_v), z(NR, 3)
    call("ndjir_neus_alpha_forward", P, N, a_fg, sdf, nrm, 3, raydir, t_fg, gain, cos_anneal)
    if Nb:
        call("ndjir_bg_alpha_forward", rb, Nb, a_bg, h0, 1, t_bg)
    call("ndjir_composite_forward", NR, N, Nb, a_fg, mask, a_bg if Nb else None, w, T)
    if Nb:
        call("ndjir_bg_color_forward", NR, Nb, w.data_ptr() + 4 * N, S, raw, 4, colbg)
    call("ndjir_volume_render_forward", NR, N, C, w, S, V, ld_v, pix, ld_v)
    # ---- fused ----
    a_fg2, a_bg2, w2, T2, pix2, colbg2 = z(P, 1), z(rb, 1), z(NR, S), z(NR, S), z(NR, ld_v), z(NR, 3)
    call("ndjir_render_segment_forward", NR, N, Nb, C, sdf, nrm, 3, raydir, t_fg, gain, cos_anneal, mask,
         h0 if Nb else None, 1, t_bg if Nb else None, raw if Nb else None, 4, V, ld_v, a_fg2, a_bg2 if Nb else None, w2, T2,
         pix2, ld_v, colbg2 if Nb else None)
    torch.cuda.synchronize()
    for name, x, y in (("alpha_fg", a_fg, a_fg2), ("alpha_bg", a_bg, a_bg2), ("w", w, w2), ("T", T, T2),
                       ("pix", pix[:, :C], pix2[:, :C]), ("colbg", colbg, colbg2)):
        assert torch.equal(x, y), name
    assert float(w.sum()) > 0
    # ---- backward ----
    dpix, dpix2, dcol = dev(rng.randn(NR, ld_v)), dev(rng.randn(NR, C2)), dev(rng.randn(NR, 3))
    dV, dV2, dw = z(P, ld_v), z(P, C2), z(NR, S)
    call("ndjir_volume_render_backward", NR, N, C, w, S, V, ld_v, dpix, ld_v, dV, ld_v, 0, dw, S)
    call("ndjir_volume_render_backward", NR, N, C2, w, S, V2, C2, dpix2, C2, dV2, C2, 0, dw, S)
    draw = z(rb, 4)
    if Nb:
        call("ndjir_bg_color_backward", NR, Nb, w.data_ptr() + 4 * N, S, raw, 4, dcol, dw.data_ptr() + 4 * N, S, draw, 4)
    da_fg, da_bg = z(P, 1), z(rb, 1)
    call("ndjir_composite_backward", NR, N, Nb, a_fg, mask, a_bg if Nb else None, T, dw, da_fg, da_bg if Nb else None)
    dsdf, dg, dh0 = z(P, 1), z(4), z(rb, 1)
    dn = dV[:, 3:6]                 # the normal columns live inside dV, as in the engine
    call("ndjir_neus_alpha_backward", P, N, da_fg, sdf, nrm, 3, raydir, t_fg, gain, cos_anneal, dsdf, dn.data_ptr(), ld_v, dg)
    if Nb:
        call("ndjir_bg_alpha_backward", rb, Nb, da_bg, h0, 1, t_bg, dh0, 1)
    dVf, dV2f, drawf, dh0f, dsdff, dgf = z(P, ld_v), z(P, C2), z(rb, 4), z(rb, 1), z(P, 1), z(4)
    dwf, daf, dabf = z(NR, S), z(P, 1), z(rb, 1)
    call("ndjir_render_segment_backward", NR, N, Nb, C, C2, sdf, nrm, 3, raydir, t_fg, gain, cos_anneal, mask,
         h0 if Nb else None, 1, t_bg if Nb else None, raw if Nb else None, 4, V, ld_v, V2, C2, w, T, dpix, ld_v, dpix2, C2,
         dcol if Nb else None, dVf, ld_v, dV2f, C2, drawf if Nb else None, 4, dh0f if Nb else None, 1, dsdff,
         dVf.data_ptr() + 12, ld_v, dgf, dwf, daf, dabf if Nb else None)
    torch.cuda.synchronize()

    def close(name, x, y, tol=1e-6):
        sc = max(float(y.abs().max()), 1e-30)
        assert float((x - y).abs().max()) <= tol * sc, (name, float((x - y).abs().max()), sc)
    close("dV", dVf[:, :C], dV[:, :C]); close("dV2", dV2f, dV2); close("dw", dwf, dw, 2e-6)
    close("dalpha_fg", daf, da_fg, 1e-5); close("dsdf", dsdff, dsdf, 1e-5); close("dgain", dgf[:1], dg[:1], 1e-4)
    if Nb:
        close("draw", drawf[:, :3], draw[:, :3]); close("dalpha_bg", dabf, da_bg, 1e-5); close("dh0", dh0f, dh0, 1e-5)
