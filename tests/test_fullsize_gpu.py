"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run these sizes in
seconds): default.yaml networks, 512^3 x 4 voxel grid, 4 x 512 rays x (128 + 32) samples; grid micro-benchmark shape
2^22 points.

  * compositing closure: sum_i w_i + T_end = 1 for every ray, weights in [0,1], transmittance non-increasing;
  * sample placement: distances sorted, inside [t_near, t_far], masks consistent with the hit counts;
  * trilinear partition of unity: a constant grid is reproduced exactly and the scatter-add conserves mass
    (sum over cells of grad_feature == sum over points of grad_output: a checksum of checksums);
  * linearity of the query in the features;
  * the train step is deterministic up to atomic order and its gradients are finite, with the gradient of the total
    loss equal to the sum of the gradients of its parts is covered at small size by test_engine_gpu.py.
"""
import numpy as np
import pytest
import torch

from ndjir_b200 import scene
from ndjir_b200 import grid_feature as GF
from ndjir_b200.config import make_conf
from ndjir_b200.engine import Engine

pytestmark = pytest.mark.gpu
MN, MX = [-1.0] * 3, [1.0] * 3


@pytest.fixture(scope="module")
def full():
    conf = make_conf("default")
    eng = Engine(conf)
    eng.params.load_reference(scene.init_params(conf, seed=313))
    eng.params.init_grid_on_device(scene.grid_shapes(conf), std=1e-3, seed=313)
    tr = conf.train
    camloc, raydir, gt = scene.make_batch(conf, step=0)
    raydir[0, :7] = -raydir[0, :7]          # a few rays that miss the box
    rnd = scene.make_randoms(conf, tr.batch_size, tr.n_rays)
    d = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return conf, eng, d(camloc), d(raydir), d(gt), {k: d(v) for k, v in rnd.items()}


def test_sample_placement_properties(full):
    conf, eng, camloc, raydir, gt, rnd = full
    x_fg, t_fg, x_bg, t_bg, mask = eng.sample_points(camloc, raydir, rnd["stratified"], rnd["background"])
    torch.cuda.synchronize()
    B, R, N1, _ = t_fg.shape
    t = t_fg[..., 0]
    assert torch.isfinite(t).all()
    assert (t[:, :, 1:] >= t[:, :, :-1]).all(), "foreground distances must be sorted"
    tn = eng.buf("t_near", B * R, 1)[:B * R].view(B, R, 1)
    tf = eng.buf("t_far", B * R, 1)[:B * R].view(B, R, 1)
    assert (t >= tn - 1e-6).all() and (t <= tf + 1e-6).all()
    assert torch.equal(t[:, :, -1:], tf), "t_fg ends with t_far (sampler.py:280)"
    m = mask.reshape(B, R)
    assert m[0, :7].sum() == 0 and m.sum() >= 0.9 * B * R
    tb = t_bg[..., 0]
    assert (tb[:, :, 1:] >= tb[:, :, :-1]).all()
    assert torch.allclose((x_bg[..., :3] ** 2).sum(-1), torch.ones_like(x_bg[..., 0]), atol=1e-4)   # unit directions
    # x_fg = o + t d
    want = camloc.view(B, 1, 1, 3) + t_fg[:, :, :-1] * raydir.view(B, R, 1, 3)
    assert torch.allclose(x_fg, want, atol=1e-6)


def test_train_step_full_size_properties(full):
    conf, eng, camloc, raydir, gt, rnd = full
    losses = eng.train_step(camloc, raydir, gt, rnd, cos_anneal_ratio=0.3, keep=True)
    torch.cuda.synchronize()
    d = eng.debug
    B, R, N, Nb, M = d["dims"]
    NR = B * R
    w, T = d["w"][:NR], d["T"][:NR]
    a_last = d["alpha_bg"][:NR * Nb].view(NR, Nb)[:, -1]
    closure = w.sum(dim=1) + T[:, -1] * (1 - a_last)          # sum w + remaining transmittance
    assert torch.allclose(closure, torch.ones_like(closure), atol=2e-5)
    assert (w >= 0).all() and (w <= 1 + 1e-6).all()
    assert (T[:, 1:] <= T[:, :-1] + 1e-6).all(), "transmittance is non-increasing (warp-scan product order: a few ulp)"
    assert (T[:, 0] == 1).all()
    # rays that miss: foreground weights are exactly zero (alpha_fg * mask, renderer.py:79)
    assert (w[:7, :N] == 0).all()
    assert torch.isfinite(losses).all()
    g = eng.params.grad
    assert torch.isfinite(g).all() and g.abs().max() > 0
    gg = eng.params.grid_grad["voxel"]
    nz = (gg != 0).any(dim=-1).sum().item()
    assert 0 < nz <= 2 * 8 * NR * N + 4 * NR * N, "only cells touched by samples (and their TV neighbours) get gradient"
    color = d["color"][:NR]
    assert torch.isfinite(color).all() and (color >= 0).all()
    # determinism up to atomic order
    l2 = eng.train_step(camloc, raydir, gt, rnd, cos_anneal_ratio=0.3)
    torch.cuda.synchronize()
    assert torch.allclose(losses, l2, rtol=1e-5, atol=1e-7)
    assert torch.allclose(g, eng.params.grad, rtol=2e-3, atol=1e-7 * float(g.abs().max()) + 1e-9)


def test_voxel_partition_of_unity_and_mass_conservation(full):
    conf, eng, *_ = full
    G, D = 512, 4
    n = 1 << 22
    gen = torch.Generator(device="cuda").manual_seed(412)
    q = torch.rand((n, 3), device="cuda", generator=gen) * 2 - 1
    const = torch.full((G, G, G, D), 0.37, device="cuda")
    out = GF.query_on_voxel(q, const, MN, MX)
    assert torch.allclose(out, torch.full_like(out, 0.37), atol=1e-6)
    del const
    F = eng.params.grid["voxel"].detach().clone().requires_grad_(True)
    go = torch.randn((n, D), device="cuda", generator=gen)
    (GF.query_on_voxel(q, F, MN, MX) * go).sum().backward()
    got = F.grad.double().sum(dim=(0, 1, 2))
    want = go.double().sum(dim=0)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-3), (got, want)


def test_query_is_linear_in_features(full):
    conf, eng, *_ = full
    n = 1 << 20
    gen = torch.Generator(device="cuda").manual_seed(7)
    q = torch.rand((n, 3), device="cuda", generator=gen) * 2 - 1
    F1 = eng.params.grid["voxel"]
    F2 = torch.randn(F1.shape, device="cuda", generator=gen) * 1e-3
    a, b = 0.7, -1.3
    lhs = GF.query_on_voxel(q, a * F1 + b * F2, MN, MX)
    rhs = a * GF.query_on_voxel(q, F1, MN, MX) + b * GF.query_on_voxel(q, F2, MN, MX)
    assert torch.allclose(lhs, rhs, atol=1e-8 + 1e-5 * float(rhs.abs().max()))
