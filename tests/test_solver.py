"""Optimizer row: oracle/cpu_solver.py (numpy restatement of python/solver.py + nnabla's
published Adam rule; parity unpinned - nnabla is absent) against closed forms on the CPU, and the fused CUDA step
(csrc/optimizer.cu through ndjir_b200/solver.py) against the oracle on the GPU."""
import math

import numpy as np
import pytest

from oracle import cpu_solver as S
from ndjir_b200.config import make_conf


def test_schedules_closed_forms():
    conf = make_conf("default")
    lw, lf = S.learning_rates(conf)
    assert lw == pytest.approx(5e-4 * 4) and lf == pytest.approx(5e-4 * 4)        # 2048 rays / 512
    wu = int(1500 * 0.015)                                                          # 22 warm-up epochs
    assert S.compute_learning_rate(conf, 0, lw) == 0.0                              # S.Adam(0) + lr*0/22 (q17)
    assert S.compute_learning_rate(conf, 11, lw) == pytest.approx(lw * 11 / wu)
    # continuous at the end of the warm-up, lr_end_ratio at the last epoch
    assert S.compute_learning_rate(conf, wu, lw) == pytest.approx(lw, rel=2e-3)
    assert S.compute_learning_rate(conf, 1500, lw) == pytest.approx(0.01 * lw, rel=1e-9, abs=1e-12)
    assert S.cos_anneal_ratio(conf, 0) == 1.0 and S.cos_anneal_ratio(conf, 225) == 1.0
    assert S.cos_anneal_ratio(conf, 112) == pytest.approx(0.5 * math.cos(math.pi * 112 / 225) + 0.5)
    assert S.light_visibility_gain(conf, 0) == 1.0 and S.light_visibility_gain(conf, 1500) == 1.0


def test_host_schedules_match_oracle():
    from ndjir_b200.solver import Solvers

    class _Eng:
        class params:
            pl_gain = 0.0
        device = "cpu"
    conf = make_conf("default", train={"sigmoid_gain_lv_end": 3})
    sv = Solvers(conf, _Eng())
    lw, lf = S.learning_rates(conf)
    assert (sv.learning_rate_weight, sv.learning_rate_feat) == (lw, lf)
    for i in (0, 1, 21, 22, 23, 100, 224, 225, 700, 1499):
        sv.update_learning_rate(i)
        assert sv.lr_weight == pytest.approx(S.compute_learning_rate(conf, i, lw), rel=1e-12, abs=0)
        assert sv.cos_anneal_ratio == pytest.approx(S.cos_anneal_ratio(conf, i), rel=1e-12)
        assert _Eng.params.pl_gain == pytest.approx(S.light_visibility_gain(conf, i), rel=1e-12)


def test_oracle_adam_first_step_and_fixed_point():
    """t = 1: m = (1-b1) g, v = (1-b2) g^2, alpha_1 = alpha sqrt(1-b2)/(1-b1), so |dw| = alpha |g|/(|g| + eps') ~ alpha;
    zero gradient and zero state leave the weights unchanged (padding entries of the flat buffer)."""
    w = np.array([1.0, -2.0, 0.5, 0.0], np.float32)
    g = np.array([0.3, -7.0, 1e-3, 0.0], np.float32)
    a = S.Adam(alpha=1e-2)
    w0 = w.copy()
    a.update({"w": (w, g)})
    np.testing.assert_allclose(w[:3], w0[:3] - 1e-2 * np.sign(g[:3]), rtol=1e-4)
    assert w[3] == 0.0


def _groups(rng, sizes):
    return {f"p{i}": ((rng.randn(n) * 0.1).astype(np.float32), np.zeros(n, np.float32)) for i, n in enumerate(sizes)}


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["default", "no_voxel"])
def test_fused_step_matches_oracle(kind):
    """Five iterations of zero_grad / weight_decay / backward / check / update (train.py:135-148) on a small engine:
    the fused device step vs the numpy oracle fed the same synthetic gradients, one of them non-finite in BOTH groups
    (update skipped, Adam's t not advanced; with no feature grid the empty feat solver never reports
    a bad gradient, so nothing is ever skipped and the non-finite entry propagates, as in the reference) and one non-finite in the weight group only (`and`: update NOT skipped)."""
    import torch
    from ndjir_b200.engine import Engine
    from ndjir_b200.solver import Solvers
    from ndjir_b200 import scene
    from test_engine_gpu import small_conf
    conf = small_conf(kind)
    eng = Engine(conf)
    eng.params.load_reference(scene.init_params(conf, seed=313, grid_std=0.05))
    ps = eng.params
    sv = Solvers(conf, eng)
    sv.set_parameters()
    rng = np.random.RandomState(5)
    weights = {"flat": (ps.data.cpu().numpy().copy(), np.zeros(ps.data.numel(), np.float32))}
    feats = {k: (v.cpu().numpy().copy().reshape(-1), np.zeros(v.numel(), np.float32)) for k, v in ps.grid.items()}
    ow, of = S.Adam(0.0), S.Adam(0.0)
    lw, lf = S.learning_rates(conf)
    ps.zero_grad()
    for it, epoch in enumerate((30, 31, 32, 33, 34)):
        sv.update_learning_rate(epoch)
        ow.set_learning_rate(S.compute_learning_rate(conf, epoch, lw))
        of.set_learning_rate(S.compute_learning_rate(conf, epoch, lf))
        gw = (rng.randn(ps.data.numel()) * 1e-2).astype(np.float32)
        gw[rng.rand(gw.size) < 0.3] = 0.0
        gf = {k: (rng.randn(v.numel()) * 1e-3).astype(np.float32) * (rng.rand(v.numel()) < 0.1) for k, v in ps.grid.items()}
        gf = {k: v.astype(np.float32) for k, v in gf.items()}
        if it == 1:                                    # both groups bad -> skipped (single-solver config: weight only)
            gw[7] = np.inf
            for v in gf.values():
                v[3] = np.nan
        if it == 3:                                    # weight group only -> NOT skipped when a feat group exists
            gw[11] = np.nan
        # device: gradients accumulate into the zeroed buffers, as train_step(zero_grad=False) would
        ps.grad += torch.from_numpy(gw).cuda()
        for k, v in gf.items():
            ps.grid_grad[k] += torch.from_numpy(v).cuda().view_as(ps.grid_grad[k])
        sv.step()

        def backward():
            weights["flat"][1][...] += gw
            for k, v in gf.items():
                feats[k][1][...] += v
        with np.errstate(invalid="ignore", over="ignore"):
            done = S.train_iteration(conf, ow, of, weights, feats, backward)
        if it == 1:     # an empty feat solver reports False, so without a grid the reference never skips
            assert done == (not feats)
        if it == 3:
            assert done
        torch.cuda.synchronize()
        assert int(sv._t.item()) == ow.t, "Adam step count"
        assert float(ps.grad.abs().max()) == 0.0 and all(float(g.abs().max()) == 0.0 for g in ps.grid_grad.values())
        dw = ps.data.cpu().numpy()
        ref = weights["flat"][0]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(dw), fin)
        scale = np.abs(ref[fin]).max()
        assert np.abs(dw[fin] - ref[fin]).max() / scale < 2e-6, f"iteration {it}: flat weights"
        for k, v in ps.grid.items():
            r = feats[k][0]
            assert np.abs(v.cpu().numpy().reshape(-1) - r).max() / np.abs(r).max() < 2e-6, f"iteration {it}: grid {k}"


@pytest.mark.gpu
def test_reference_call_order_equals_fused_step():
    """zero_grad(); weight_decay(); [backward]; check_inf_or_nan_grad(); update() - the reference's own order - gives
    the same parameters as the fused step()."""
    import torch
    from ndjir_b200.engine import Engine
    from ndjir_b200.solver import Solvers
    from ndjir_b200 import scene
    from test_engine_gpu import small_conf
    conf = small_conf("default")
    out = []
    for fused in (True, False):
        eng = Engine(conf)
        eng.params.load_reference(scene.init_params(conf, seed=313, grid_std=0.05))
        ps = eng.params
        sv = Solvers(conf, eng)
        sv.set_parameters()
        gen = torch.Generator(device="cuda").manual_seed(3)
        ps.zero_grad()
        for epoch in (40, 41, 42):
            sv.update_learning_rate(epoch)
            gw = torch.randn(ps.data.numel(), device="cuda", generator=gen) * 1e-2
            gf = {k: torch.randn(v.shape, device="cuda", generator=gen) * 1e-3 for k, v in sorted(ps.grid.items())}
            if fused:
                ps.grad += gw
                for k in gf:
                    ps.grid_grad[k] += gf[k]
                sv.step()
            else:
                sv.zero_grad(); sv.weight_decay(); sv.clip_grad_by_norm()
                ps.grad += gw
                for k in gf:
                    ps.grid_grad[k] += gf[k]
                assert not sv.check_inf_or_nan_grad()
                sv.update()
        out.append((ps.data.clone(), {k: v.clone() for k, v in ps.grid.items()}))
    (a, ga), (b, gb) = out
    assert (a - b).abs().max().item() <= 1e-6 * a.abs().max().item()
    for k in ga:
        assert (ga[k] - gb[k]).abs().max().item() <= 1e-6 * ga[k].abs().max().item()


@pytest.mark.gpu
def test_nan_loss_skips_the_fused_step_without_a_feature_grid():
    """voxel.type = none: the feat solver is empty, so the gradient `and` can never trip (solver.py:67-69); the loop's
    second guard - `if np.any(np.isnan(loss.d)): continue`, train.py:144-146 - must skip the iteration instead of letting
    Adam write NaN into every MLP weight.  A finite loss with the same gradients is NOT skipped."""
    import torch
    from ndjir_b200.engine import Engine
    from ndjir_b200.solver import Solvers
    from ndjir_b200 import scene
    from test_engine_gpu import small_conf
    conf = small_conf("no_voxel")
    eng = Engine(conf)
    eng.params.load_reference(scene.init_params(conf, seed=313))
    ps = eng.params
    sv = Solvers(conf, eng)
    sv.set_parameters()
    sv.update_learning_rate(30)
    before = ps.data.clone()
    ps.zero_grad()
    ps.grad += float("nan")
    sv.step(loss=torch.tensor([float("nan")], device="cuda"))
    torch.cuda.synchronize()
    assert int(sv._t.item()) == 0, "skipped iterations do not advance Adam's count"
    assert torch.equal(ps.data, before), "a NaN loss skips the update"
    assert float(ps.grad.abs().max()) == 0.0, "the gradient is still zeroed for the next iteration"
    ps.grad += 1e-3
    sv.step(loss=torch.tensor([0.5], device="cuda"))
    torch.cuda.synchronize()
    assert int(sv._t.item()) == 1 and not torch.equal(ps.data, before)
    assert bool(torch.isfinite(ps.data).all())
