"""CPU: the oracle of the nnabla-composed half of the path (oracle/cpu_render.py, oracle/cpu_solver.py) against golden
vectors produced by EXECUTING the reference's own python/sampler.py, network.py, renderer.py, specular_brdf.py,
loss.py and solver.py (tests/golden/make_golden.py::golden_render through tests/golden/nnabla_standin.py).

Both sides run in float64 on the same seeded inputs (tests/golden/cases.py), so the bar is tight: 1e-9 relative
(max-norm per tensor) for every output of sample_points, pb_render and total_loss and for every parameter gradient -
any difference in op order, clipping, concatenation order, parameter naming / layout or loss normalisation shows up
as 1e-3 or worse.  This is what pins the oracle at the nnabla boundary."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import cpu_render as CR
from oracle import cpu_solver as CS

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


def rel(a, b):
    a = np.asarray(a.detach().numpy() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module", params=list(cases.CASES))
def case(request):
    name = request.param
    g = np.load(os.path.join(GOLD, f"render_{name}.npz"))
    conf, P, camloc, raydir, color_gt, rnd, cos_anneal = cases.build_case(name)
    model = CR.Model(conf, P, dtype=torch.float64)
    return dict(name=name, g=g, conf=conf, P=P, camloc=camloc, raydir=raydir, color_gt=color_gt, rnd=rnd,
                cos_anneal=cos_anneal, model=model, obj_mask=cases.obj_mask_of(name))


def test_sample_points_equals_reference(case):
    g, m = case["g"], case["model"]
    out = CR.sample_points(m, case["camloc"], case["raydir"], case["rnd"]["stratified"], case["rnd"]["background"])
    for k, v in zip(("x_fg", "t_fg", "x_bg", "t_bg", "mask"), out):
        want = g[f"samples.{k}"]
        if k == "mask":
            assert np.array_equal(v.numpy().reshape(want.shape), want), "hit mask"
        else:
            assert rel(v, want) < TOL, (k, rel(v, want))
    r = case["conf"].renderer
    assert g["samples.x_fg"].shape[2] == r.n_samples0 + r.n_upsamples * r.n_samples1 == 128
    assert g["samples.x_bg"].shape[2] == 32


def test_pb_render_and_losses_equal_reference(case):
    g, m = case["g"], case["model"]
    samples = [torch.as_tensor(g[f"samples.{k}"]) for k in ("x_fg", "t_fg", "x_bg", "t_bg", "mask")]
    losses, res, _ = CR.total_loss(m, case["camloc"], case["raydir"], case["color_gt"], case["cos_anneal"], case["rnd"],
                                   return_all=True, samples=samples, obj_mask=case["obj_mask"])
    for k in ("color_pixel", "sdf_x_fg", "grad_x_fg", "alpha_fg", "trans_fg", "base_color", "base_color_ptb",
              "roughness", "specular_reflectance", "std_roughness", "std_specular_reflectance"):
        assert rel(res[k], g[f"render.{k}"]) < TOL, (k, rel(res[k], g[f"render.{k}"]))
    for k in ("loss", "loss_rgb", "loss_eikonal", "loss_tv", "loss_mask", "prior_base_color", "prior_roughness",
              "prior_specular_reflectance", "reg_std_roughness", "reg_std_specular_reflectance"):
        want = float(g[f"loss.{k}"])
        got = float(losses[k].detach())
        assert abs(got - want) <= TOL * max(abs(want), 1e-12), (k, got, want)
    if case["obj_mask"] is None:
        assert float(g["loss.loss_mask"]) == 0.0
    else:       # the mask term of loss.py:108-116 on the reference's own obj_mask_pred (renderer.py:183-185)
        assert float(g["loss.loss_mask"]) > 0.0
        assert rel((res["alpha_fg"] * res["trans_fg"]).sum(dim=2), g["render.obj_mask_pred"]) < TOL


def test_gradients_equal_reference(case):
    """loss.backward() of the reference (every MLP, the gain, the grid tables) against the oracle's autograd."""
    g, m = case["g"], case["model"]
    _, grads = CR.train_step(m, case["camloc"], case["raydir"], case["color_gt"], case["cos_anneal"], case["rnd"],
                             obj_mask=case["obj_mask"])
    seen = 0
    for k, gr in grads.items():
        gr = gr.numpy()
        if f"grad.{k}" in g.files:
            want = g[f"grad.{k}"]
            scale = max(np.abs(want).max(), 1e-300)
            assert np.abs(gr.reshape(want.shape) - want).max() / scale < 1e-8, k
            seen += 1
        else:
            assert f"gradS.{k}.norm" in g.files, f"gradient {k} missing from the golden file"
            flat = gr.reshape(-1)
            scale = max(np.abs(g[f"gradS.{k}.topval"]).max(), 1e-300)
            assert abs(np.linalg.norm(flat) - float(g[f"gradS.{k}.norm"])) <= 1e-8 * float(g[f"gradS.{k}.norm"]) + 1e-300, k
            assert np.abs(flat[g[f"gradS.{k}.pos"]] - g[f"gradS.{k}.val"]).max() / scale < 1e-8, k
            assert np.abs(flat[g[f"gradS.{k}.top"]] - g[f"gradS.{k}.topval"]).max() / scale < 1e-8, k
            seen += 1
    n_gold = len([f for f in g.files if f.startswith("grad.")]) + len([f for f in g.files if f.endswith(".norm")])
    assert seen == n_gold == len(grads), (seen, n_gold, len(grads))


def test_full_case_has_default_widths():
    conf = cases.case_conf("full_default")
    from ndjir_b200.scene import network_dims
    d = network_dims(conf)
    assert d["geo"][3] == (256, 213) and d["geo"][4] == (256, 256) and d["geo"][-1] == (256, 257)
    assert d["bc"][0] == (259, 256) and d["sv"][0] == (301, 128) and d["pl"][0] == (290, 256)


def test_parameter_names_are_the_references():
    """The name table the golden generator fed the reference's network.py through (it fails on any name the reference
    asks for that the table lacks, and on any table entry the reference never asks for)."""
    from ndjir_b200 import nnabla_names
    conf = cases.case_conf("full_default")
    names = [n for n, _ in nnabla_names.parameter_names(conf)]
    assert "geometric-network/affine-last/affine/W" in names and "geometric-network/voxel_feature/F" in names
    assert "roughness-network/affine--1/affine/W" in names and "roughness-network/affine-03/affine/b" in names   # q16
    assert "specular-reflectance-network/affine-01/affine/W" in names
    assert "background-network/lighting-network/affine-01/affine/W" in names
    assert len(names) == len(set(names))
    P = cases.build_case("full_default")[1]
    back = nnabla_names.from_nnabla(conf, nnabla_names.to_nnabla(conf, P))
    for net in ("geo", "ro", "bg1"):
        for (W, b), (W2, b2) in zip(P[net], back[net]):
            assert W is W2 and b is b2


# ----------------------------------------------------------------------------------------------------
# solver.py: schedules and the iteration order of train.py:135-148
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,over", [("default", {}), ("short", {"epoch": 40, "warmup_term_ratio": 0.1,
                                                                  "sigmoid_gain_lv_end": 3})])
def test_solver_schedules_and_adam_equal_reference(tag, over):
    g = np.load(os.path.join(GOLD, "solver.npz"))
    conf = cases.case_conf("small_default")
    for k, v in over.items():
        setattr(conf.train, k, v)
    lw, lf = CS.learning_rates(conf)
    for j, i in enumerate(g[f"{tag}.iters"]):
        i = int(i)
        assert abs(CS.compute_learning_rate(conf, i, lw) - g[f"{tag}.lr"][j, 0]) <= 1e-12 * max(lw, 1e-30)
        assert abs(CS.compute_learning_rate(conf, i, lf) - g[f"{tag}.lr"][j, 1]) <= 1e-12 * max(lf, 1e-30)
        assert abs(CS.cos_anneal_ratio(conf, i) - g[f"{tag}.cos_anneal_ratio"][j]) <= 1e-12
        assert abs(CS.light_visibility_gain(conf, i) - g[f"{tag}.pl_gain"][j]) <= 1e-12
    # three iterations (zero_grad, weight_decay, backward, check, update) in float32 against the float64 reference run
    names = ["net/affine/W", "net/affine/b", "geo/voxel_feature/F"]
    w0 = {"net/affine/W": np.linspace(-1, 1, 12).reshape(3, 4), "net/affine/b": np.zeros(4),
          "geo/voxel_feature/F": np.linspace(0.1, 0.5, 8).reshape(2, 4)}
    params = {n: (w0[n].astype(np.float32), np.zeros(w0[n].shape, np.float32)) for n in names}
    weights = {n: params[n] for n in names[:2]}
    feats = {names[2]: params[names[2]]}
    sw, sf = CS.Adam(), CS.Adam()
    E = conf.train.epoch
    sw.set_learning_rate(CS.compute_learning_rate(conf, E // 4, lw))
    sf.set_learning_rate(CS.compute_learning_rate(conf, E // 4, lf))
    for it in range(3):
        def backward():
            for n in names:
                params[n][1][...] += g[f"{tag}.it{it}.g.{n}"].astype(np.float32)
        assert CS.train_iteration(conf, sw, sf, weights, feats, backward)
        for n in names:
            np.testing.assert_allclose(params[n][0], g[f"{tag}.it{it}.w.{n}"], rtol=2e-5, atol=2e-7)
