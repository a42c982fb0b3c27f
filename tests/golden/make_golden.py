"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON where it lies under /root/reference.

Nothing is copied: functions are pulled out of the reference files with `ast` at run time and executed here.
  * intersection / direction sampling: the numpy oracles inside the reference's tests
    (python/intersection/test/test_ray_aabb_intersection.py:24-110, test_ray_sphere_intersection.py:25-76,
    python/sampler/test_sampler.py:23-70), driven with the reference tests' seeds and parametrisations.
  * grid families: the reference's pure-op "composite" statements
    (python/grid_feature/{voxel,triplane,triline,lanczos_voxel,cosine_voxel,cosine_triplane,cosine_triline}_feature_composite.py,
    total_variation_loss{,_on_triplane,_on_triline}_composite.py) executed through a small torch(float64)-backed
    stand-in for `nnabla.functions` (only the dozen ops those files use); first- and second-order gradients
    come from torch autograd of that same reference code, mirroring what the reference tests do with nnabla
    (python/grid_feature/test/test_voxel_feature.py:25-150).
nnabla itself is not installable in this image (no network), which is why the stand-in exists.

Run here (needs /root/reference):  python tests/golden/make_golden.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NDJIR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def extract_functions(path, names, env):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), env)
    missing = [n for n in names if n not in env]
    assert not missing, (path, missing)
    return env


# ----------------------------------------------------------------------------------------------------
# numpy oracles from the reference tests
# ----------------------------------------------------------------------------------------------------
def golden_intersection():
    env = {"np": np}
    extract_functions(f"{REF}/python/intersection/test/test_ray_aabb_intersection.py",
                      ["ray_aabb_intersection_python"], env)
    extract_functions(f"{REF}/python/intersection/test/test_ray_sphere_intersection.py",
                      ["ray_sphere_intersection_python"], env)
    extract_functions(f"{REF}/python/intersection/ray_sphere_intersection.py", ["sample_inside_sphere"], env)
    out = {}
    seed, B, R = 412, 2, 3
    # test_ray_aabb_intersection.py:113-131
    for k, (radius, size) in enumerate([(3, 1), (3, 1.5), (1, 2)]):
        rng = np.random.RandomState(seed)
        camloc = rng.randn(B, 3)
        camloc /= np.linalg.norm(camloc, ord=2, axis=-1, keepdims=True)
        camloc *= radius
        raydir = rng.rand(B, R, 3) * size * 2 - size
        raydir = raydir - camloc.reshape((B, 1, 3))
        raydir /= np.linalg.norm(raydir, ord=2, axis=-1, keepdims=True)
        camloc, raydir = camloc.astype(np.float32), raydir.astype(np.float32)
        tn, tf, nh = env["ray_aabb_intersection_python"](camloc, raydir, size)
        out[f"aabb{k}_camloc"], out[f"aabb{k}_raydir"], out[f"aabb{k}_size"] = camloc, raydir, np.float32(size)
        out[f"aabb{k}_t_near"], out[f"aabb{k}_t_far"], out[f"aabb{k}_n_hits"] = (
            np.asarray(tn, dtype=np.float64).reshape(-1), np.asarray(tf, dtype=np.float64).reshape(-1),
            np.asarray(nh, dtype=np.float64).reshape(-1))
    # test_ray_sphere_intersection.py:79-111
    for k, (radius, ratio) in enumerate([(1, 2), (1.5, 2), (1.0, 0.5)]):
        rng = np.random.RandomState(seed)
        camloc = rng.randn(B, 3)
        camloc /= np.linalg.norm(camloc, ord=2, axis=-1, keepdims=True)
        camloc *= (radius * ratio)
        raydir = env["sample_inside_sphere"](B, R, radius, rng) - camloc.reshape((B, 1, 3))
        raydir /= np.linalg.norm(raydir, ord=2, axis=-1, keepdims=True)
        camloc, raydir = camloc.astype(np.float32), raydir.astype(np.float32)
        tn, tf, nh = env["ray_sphere_intersection_python"](camloc, raydir, radius)
        out[f"sphere{k}_camloc"], out[f"sphere{k}_raydir"], out[f"sphere{k}_radius"] = camloc, raydir, np.float32(radius)
        out[f"sphere{k}_t_near"], out[f"sphere{k}_t_far"], out[f"sphere{k}_n_hits"] = (
            np.asarray(tn, dtype=np.float64).reshape(-1), np.asarray(tf, dtype=np.float64).reshape(-1),
            np.asarray(nh, dtype=np.float64).reshape(-1))
    np.savez(os.path.join(OUT, "intersection.npz"), **out)
    return out


def golden_directions():
    env = {"np": np}
    extract_functions(f"{REF}/python/sampler/test_sampler.py", ["sample_directions_numpy"], env)
    out = {}
    seed, k = 412, 0
    # test_sampler.py:73-111 (eps does not enter the numpy oracle)
    for (B, R) in [(1, 1), (2, 4)]:
        for n_thetas in [1, 4]:
            for typ in ["uniform", "importance"]:
                rng = np.random.RandomState(seed)
                normal = rng.randn(B, R, 3).astype(np.float32)
                normal = normal / np.linalg.norm(normal, ord=2, axis=-1, keepdims=True)
                cdf_the = rng.rand(B, R, n_thetas).astype(np.float32)
                cdf_phi = rng.rand(B, R, 2 * n_thetas).astype(np.float32)
                out[f"dir{k}_normal"], out[f"dir{k}_cdf_the"], out[f"dir{k}_cdf_phi"] = normal, cdf_the, cdf_phi
                if typ == "uniform":
                    dirs = env["sample_directions_numpy"](normal, cdf_the, cdf_phi)
                else:
                    alpha = rng.randn(B, R, 1).astype(np.float32)
                    out[f"dir{k}_alpha"] = alpha
                    dirs = env["sample_directions_numpy"](normal, cdf_the, cdf_phi, alpha)
                out[f"dir{k}_dirs"] = np.asarray(dirs, dtype=np.float64)
                k += 1
    out["n_cases"] = np.int64(k)
    np.savez(os.path.join(OUT, "directions.npz"), **out)
    return out


# ----------------------------------------------------------------------------------------------------
# torch(float64)-backed stand-in for the handful of nnabla.functions ops used by the composite files
# ----------------------------------------------------------------------------------------------------
class V(torch.Tensor):
    """Tensor with nnabla.Variable's `.apply(need_grad=...)` and float-valued advanced indexing."""

    def apply(self, need_grad=True, **kw):
        return self if need_grad else self.detach().as_subclass(V)

    def __getitem__(self, idx):
        def fix(i):
            if isinstance(i, torch.Tensor) and i.is_floating_point():
                return i.detach().as_subclass(torch.Tensor).long()
            return i
        idx = tuple(fix(i) for i in idx) if isinstance(idx, tuple) else fix(idx)
        return super().__getitem__(idx)


def as_v(x, requires_grad=False):
    t = torch.as_tensor(np.asarray(x), dtype=torch.float64).clone()
    t.requires_grad_(requires_grad)
    return t.as_subclass(V)


def make_F():
    F = types.SimpleNamespace()
    F.floor = lambda x: torch.floor(x)
    F.constant = lambda val, shape: torch.full(tuple(shape), float(val), dtype=torch.float64).as_subclass(V)
    F.reshape = lambda x, shape, inplace=True: x.reshape(tuple(shape))
    F.concatenate = lambda *xs, axis=-1: torch.cat([x for x in xs], dim=axis)
    F.stack = lambda *xs, axis=0: torch.stack([x for x in xs], dim=axis)
    F.clip_by_value = lambda x, a, b: torch.clamp(x, min=a, max=b)
    F.sinc = lambda x: torch.sinc(x / np.pi)           # nnabla sinc is sin(x)/x, torch's is normalised
    F.cos = lambda x: torch.cos(x)
    F.transpose = lambda x, axes: x.permute(*axes)
    F.greater_equal_scalar = lambda x, v: (x >= v).to(torch.float64)
    F.less_equal_scalar = lambda x, v: (x <= v).to(torch.float64)
    return F


def load_composite(relpath, names):
    env = {"F": make_F(), "np": np}
    src = open(f"{REF}/python/grid_feature/{relpath}").read()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    exec(compile(ast.Module(body=body, type_ignores=[]), relpath, "exec"), env)
    return [env[n] for n in names]


def run_query_family(fn, query, feature, mn, mx, rng, second_order=True):
    """Forward, first-order and second-order gradients of a composite, as the reference tests take them."""
    q = as_v(query, True)
    f = as_v(feature, True)
    out = fn(q, f, mn, mx)
    res = {"output": out.detach().numpy()}
    go = as_v(rng.randn(*out.shape).astype(np.float32), True)
    gq, gf = torch.autograd.grad(out, [q, f], grad_outputs=go, create_graph=True, allow_unused=True)
    res["grad_output"] = go.detach().numpy()
    res["grad_query"] = gq.detach().numpy()
    res["grad_feature"] = gf.detach().numpy()
    if second_order:
        gg = as_v(rng.randn(*q.shape).astype(np.float32))
        phi = (gq * gg).sum()
        ggo, gqgq, gqgf = torch.autograd.grad(phi, [go, q, f], allow_unused=True)
        res["grad_grad_query"] = gg.detach().numpy()
        res["gq_ggo"] = ggo.detach().numpy()
        res["gq_gf"] = gqgf.detach().numpy()
        res["gq_gq"] = (gqgq.detach().numpy() if gqgq is not None else np.zeros(q.shape))
    return {k: np.asarray(v, dtype=np.float64) for k, v in res.items()}


def golden_grids():
    mn, mx = -1.0, 1.0
    out = {}
    # reference parametrisation: seed 412, batch in {2,16}, G in {2,8}, D=4, features randn*0.01
    # (python/grid_feature/test/test_voxel_feature.py:25-45); queries are kept strictly inside the box
    # because the composites do not clamp the upper corner (voxel_feature_composite.py:26-27).
    (voxel,) = load_composite("voxel_feature_composite.py", ["query_on_voxel"])
    (triplane,) = load_composite("triplane_feature_composite.py", ["query_on_triplane"])
    (triline,) = load_composite("triline_feature_composite.py", ["query_on_triline"])
    (lanczos,) = load_composite("lanczos_voxel_feature_composite.py", ["query_on_voxel"])
    (tvv,) = load_composite("total_variation_loss_composite.py", ["tv_loss_on_voxel"])
    (tvp,) = load_composite("total_variation_loss_on_triplane_composite.py", ["tv_loss_on_triplane"])
    (tvl,) = load_composite("total_variation_loss_on_triline_composite.py", ["tv_loss_on_triline"])
    k = 0
    for B in (2, 16):
        for G in (2, 8):
            D = 4
            rng = np.random.RandomState(412)
            query = (rng.rand(B, 3).astype(np.float32) * 1.98 - 0.99).astype(np.float32)
            cases = {
                "voxel": (voxel, (rng.randn(G, G, G, D) * 0.01).astype(np.float32)),
                "triplane": (triplane, (rng.randn(3, G, G, D) * 0.01).astype(np.float32)),
                "triline": (triline, (rng.randn(3, G, D) * 0.01).astype(np.float32)),
                "lanczos_voxel": (lanczos, (rng.randn(G, G, G, D) * 0.01).astype(np.float32)),
            }
            for name, (fn, feat) in cases.items():
                res = run_query_family(fn, query, feat, mn, mx, rng)
                out[f"{name}{k}_query"], out[f"{name}{k}_feature"] = query, feat
                for kk, vv in res.items():
                    out[f"{name}{k}_{kk}"] = vv
            for name, fn, feat in (("tv_voxel", tvv, cases["voxel"][1]), ("tv_triplane", tvp, cases["triplane"][1]),
                                   ("tv_triline", tvl, cases["triline"][1])):
                for sym in (False, True):
                    q = as_v(query)
                    f = as_v(feat, True)
                    o = fn(q, f, mn, mx, sym)
                    go = as_v(rng.randn(*o.shape).astype(np.float32))
                    (gf,) = torch.autograd.grad(o, [f], grad_outputs=go)
                    tag = f"{name}{k}_sym{int(sym)}"
                    out[f"{tag}_output"] = o.detach().numpy().astype(np.float64)
                    out[f"{tag}_grad_output"] = go.detach().numpy().astype(np.float64)
                    out[f"{tag}_grad_feature"] = gf.detach().numpy().astype(np.float64)
            k += 1
    out["n_cases"] = np.int64(k)
    # cosine families (python/grid_feature/cosine_{voxel,triplane,triline}_feature_composite.py), same parametrisation;
    # a generator of their own so that the entries above keep their values
    (cvoxel,) = load_composite("cosine_voxel_feature_composite.py", ["query_on_voxel"])
    (ctriplane,) = load_composite("cosine_triplane_feature_composite.py", ["query_on_triplane"])
    (ctriline,) = load_composite("cosine_triline_feature_composite.py", ["query_on_triline"])
    (ltriplane,) = load_composite("lanczos_triplane_feature_composite.py", ["query_on_triplane"])
    (ltriline,) = load_composite("lanczos_triline_feature_composite.py", ["query_on_triline"])
    k = 0
    for B in (2, 16):
        for G in (2, 8):
            D = 4
            rng = np.random.RandomState(412)
            query = (rng.rand(B, 3).astype(np.float32) * 1.98 - 0.99).astype(np.float32)
            for name, fn, shape in (("cosine_voxel", cvoxel, (G, G, G, D)), ("cosine_triplane", ctriplane, (3, G, G, D)),
                                    ("cosine_triline", ctriline, (3, G, D)), ("lanczos_triplane", ltriplane, (3, G, G, D)),
                                    ("lanczos_triline", ltriline, (3, G, D))):
                feat = (rng.randn(*shape) * 0.01).astype(np.float32)
                res = run_query_family(fn, query, feat, mn, mx, rng)
                out[f"{name}{k}_query"], out[f"{name}{k}_feature"] = query, feat
                for kk, vv in res.items():
                    out[f"{name}{k}_{kk}"] = vv
            k += 1
    np.savez_compressed(os.path.join(OUT, "grids.npz"), **out)
    return out


# ----------------------------------------------------------------------------------------------------
# the nnabla-composed part of the path: the reference's sampler.py / network.py / renderer.py / specular_brdf.py /
# loss.py / solver.py EXECUTED through tests/golden/nnabla_standin.py on the seeded cases of tests/golden/cases.py
# ----------------------------------------------------------------------------------------------------
def golden_render():
    sys.path.insert(0, OUT)
    import cases
    import nnabla_standin as S
    from ndjir_b200 import nnabla_names
    mods = S.load_reference_modules()
    done = {}
    only = os.environ.get("NDJIR_GOLDEN_ONLY")        # regenerate one case without touching the other fixtures
    for name in cases.CASES:
        if only and name != only:
            continue
        conf, P, camloc, raydir, color_gt, rnd, cos_anneal = cases.build_case(name)
        S.set_parameters(nnabla_names.to_nnabla(conf, P))
        r, tr = conf.renderer, conf.train
        S.set_randoms(
            {r.stratified_sample_seed: [rnd["stratified"]], r.background_sample_seed: [rnd["background"]],
             r.diffuse_cdf_the_seed: [rnd["diffuse_cdf_the"]], r.diffuse_cdf_phi_seed: [rnd["diffuse_cdf_phi"]],
             r.specular_cdf_the_seed: [rnd["specular_cdf_the"]], r.specular_cdf_phi_seed: [rnd["specular_cdf_phi"]]},
            {tr.base_color_perturb_seed: [rnd["perturb"]]})
        cam, ray, gt = S.V(camloc), S.V(raydir), S.V(color_gt)
        car = S.V(np.asarray([cos_anneal]))
        out = {}
        # sample_points (sampler.py:311-314) on its own, then pb_render (renderer.py:32) on those samples
        x_fg, t_fg, x_bg, t_bg, mask = mods["sampler"].sample_points(cam, ray, S.V(rnd["stratified"]),
                                                                     S.V(rnd["background"]), conf)
        for k, v in dict(x_fg=x_fg, t_fg=t_fg, x_bg=x_bg, t_bg=t_bg, mask=mask).items():
            out[f"samples.{k}"] = v.detach().numpy().copy()
        x_fg.apply(need_grad=True)
        res = mods["renderer"].pb_render(x_fg, t_fg, x_bg, t_bg, cam, ray, mask, car, conf)
        for k, v in res.items():
            out[f"render.{k}"] = v.detach().numpy().copy()
        # total_loss (loss.py:27) forward + backward (train.py:135-140): runs sample_points + pb_render again inside
        for p in S.get_parameters().values():
            p.grad = None
        om = cases.obj_mask_of(name)
        losses = mods["loss"].total_loss(cam, ray, gt, S.V(om) if om is not None else None, car, conf)
        for k, v in losses.items():
            out[f"loss.{k}"] = np.float64(v.detach().numpy())
        losses["loss"].backward()
        asked = set(S._STATE.asked)
        table = dict(nnabla_names.parameter_names(conf))
        missing = [n for n in table if n not in asked]
        assert not missing, f"names in nnabla_names the reference never asked for: {missing}"
        for nn_name, p in S.get_parameters().items():
            key = table[nn_name]
            okey = ({"geo_gain": "geo_gain", "pl_gain": None}.get(key[0]) if key[0] in ("geo_gain", "pl_gain")
                    else f"grid.{key[1]}" if key[0] == "grid" else f"{key[0]}.{key[2]}{key[1]}")
            if okey is None:
                continue
            g = p.grad.detach().numpy() if p.grad is not None else np.zeros(tuple(p.shape))
            if g.size > cases.BIG:
                for kk, vv in cases.sampled_view(g, okey).items():
                    out[f"gradS.{okey}.{kk}"] = vv
            else:
                out[f"grad.{okey}"] = g
        np.savez_compressed(os.path.join(OUT, f"render_{name}.npz"), **out)
        done[name] = len(out)
    if only:
        return done
    # solver.py: schedules + one Solvers iteration in the order of train.py:135-148 over a tiny parameter set
    out = {}
    conf = cases.case_conf("small_default")
    Sol = mods["solver"].Solvers
    for tag, over in (("default", {}), ("short", {"epoch": 40, "warmup_term_ratio": 0.1, "sigmoid_gain_lv_end": 3})):
        for k, v in over.items():
            setattr(conf.train, k, v)
        S._STATE.strict = False
        S.set_parameters({"net/affine/W": np.linspace(-1, 1, 12).reshape(3, 4), "net/affine/b": np.zeros(4),
                          "geo/voxel_feature/F": np.linspace(0.1, 0.5, 8).reshape(2, 4)},
                         trainable=lambda n: True)
        sol = Sol(conf)
        sol.set_parameters()
        E = conf.train.epoch
        its = sorted(set([0, 1, 2, 5, E // 10, E // 4, E // 2, E - 1]))
        lrs, car_, gain_ = [], [], []
        for i in its:
            sol.update_learning_rate(i)
            lrs.append([sol.solver_weight.learning_rate(), sol.solver_feat.learning_rate()])
            car_.append(float(S._STATE.params["cos_anneal_ratio"].detach()))
            gain_.append(float(S._STATE.params["photogrammetric-light-network/gain"].detach()))
        out[f"{tag}.iters"], out[f"{tag}.lr"] = np.asarray(its), np.asarray(lrs)
        out[f"{tag}.cos_anneal_ratio"], out[f"{tag}.pl_gain"] = np.asarray(car_), np.asarray(gain_)
        # three iterations: zero_grad, weight_decay, backward (accumulate a fixed gradient), check, update
        sol.update_learning_rate(E // 4)
        rng = np.random.RandomState(5)
        names = ["net/affine/W", "net/affine/b", "geo/voxel_feature/F"]
        for it in range(3):
            sol.zero_grad()
            sol.weight_decay()
            for n in names:
                p = S._STATE.params[n]
                g = rng.randn(*p.shape)
                out[f"{tag}.it{it}.g.{n}"] = g
                p.grad = p.grad + torch.as_tensor(g)
            assert not sol.check_inf_or_nan_grad()
            sol.update()
            for n in names:
                out[f"{tag}.it{it}.w.{n}"] = S._STATE.params[n].detach().numpy().copy()
        S._STATE.strict = True
    np.savez_compressed(os.path.join(OUT, "solver.npz"), **out)
    done["solver"] = len(out)
    return done


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: golden vectors can only be regenerated where the reference is mounted")
    if not os.environ.get("NDJIR_GOLDEN_ONLY"):      # NDJIR_GOLDEN_ONLY=<render case>: write that fixture alone
        a = golden_intersection(); print("intersection.npz", len(a))
        b = golden_directions(); print("directions.npz", len(b))
        c = golden_grids(); print("grids.npz", len(c))
    print("render_*.npz / solver.npz", golden_render())
