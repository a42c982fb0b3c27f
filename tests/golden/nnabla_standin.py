"""A torch-backed stand-in for the slice of `nnabla` that the reference's per-ray path uses, so that the reference's
OWN files - python/sampler.py, network.py, renderer.py, specular_brdf.py, loss.py, solver.py - can be imported and
EXECUTED where they lie under /root/reference (nnabla 1.29 itself is not installable in this image: no network, not
in the wheelhouse).  Nothing of the reference is copied: `load_reference_modules()` puts /root/reference/python on
sys.path behind fake `nnabla*` modules and imports the files; tests/golden/make_golden.py then calls
`sample_points`, `pb_render` and `total_loss` from those modules and stores the outputs as golden vectors.

Semantics implemented here come from nnabla's public documentation (v1.29):
  PF.affine                 y = x W + b, W (in, out) under scope `affine`, parameters `W`, `b`
  F.softplus(x, beta)       log(1 + exp(beta x)) / beta
  F.cumprod(exclusive=True) shifted product starting at 1;  F.cumsum inclusive
  F.searchsorted            right=False: first i with seq[i] >= v
  F.gather(batch_dims=2)    out[b, r, m, ...] = x[b, r, idx[b, r, m], ...]
  F.sort                    ascending values
  F.clip_by_value(x, a, b)  minimum2(maximum2(x, a), b); gradient passes inside the range
  F.concatenate             default axis = last
  nn.grad(outputs, inputs)  gradient graph seeded with ones, differentiable again
  F.rand / F.randn(seed=s)  -> the explicit random tensors registered with `set_randoms` (nnabla's RNG streams cannot
                            be reproduced without nnabla; every consumer gets the same arrays)
The native CUDA extension modules the reference imports (`*_cuda`) are replaced by the CPU oracle of the native
boundary (oracle/cpu_ref.py, oracle/cpu_render.py grid queries), which is pinned separately on the reference's kernels
and composites (tests/test_oracle_golden.py, tests/test_native_gpu.py).
One deliberate choice: SURVEY q14 - `F.gather` with an index one past the end (idx == N_ts - 1 on the N_ts - 1
weights) is undefined in nnabla; the stand-in clamps the index, as oracle/cpu_render.py documents.
"""
import contextlib
import importlib
import math
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

REF = os.environ.get("NDJIR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT = torch.float64
_STATE = types.SimpleNamespace(scope=[], params=OrderedDict(), randoms={}, asked=[], strict=True)


# ----------------------------------------------------------------------------------------------------
# Variable / NdArray
# ----------------------------------------------------------------------------------------------------
class Variable(torch.Tensor):
    @staticmethod
    def from_numpy_array(a, need_grad=None):
        return V(a)

    def apply(self, need_grad=None, persistent=None, **kw):
        if need_grad is True:
            if self.is_leaf:
                self.requires_grad_(True)
            return self
        if need_grad is False:
            return self.detach().as_subclass(Variable)
        return self

    @property
    def d(self):
        return self.detach().numpy()

    @d.setter
    def d(self, v):
        with torch.no_grad():
            self.copy_(torch.as_tensor(np.asarray(v), dtype=self.dtype).reshape(self.shape))

    def copy_from(self, src):
        with torch.no_grad():
            self.copy_(src.detach().reshape(self.shape))     # NdArray.copy_from copies by element count


def V(a, requires_grad=False):
    t = a.detach().clone().to(DT) if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=DT).clone()
    t = t.as_subclass(Variable)
    if requires_grad:
        t.requires_grad_(True)
    return t


def W_(t):
    return t.as_subclass(Variable) if torch.is_tensor(t) else t


class _Out:
    """output slot of a PythonFunction: reset_shape allocates, `.data.copy_from` fills"""

    def __init__(self):
        self.data = None

    def reset_shape(self, shape, force):
        self.data = torch.zeros(tuple(shape), dtype=DT).as_subclass(Variable)


class PythonFunction:
    def __init__(self, ctx=None):
        self.ctx = ctx

    def __call__(self, *inputs):
        outs = [_Out() for _ in range(self.min_outputs())]
        self.setup_impl(list(inputs), outs)
        self.forward_impl(list(inputs), outs)
        res = [o.data for o in outs]
        return res[0] if len(res) == 1 else tuple(res)


# ----------------------------------------------------------------------------------------------------
# nn.*
# ----------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def parameter_scope(name, scope=None):
    parts = [p for p in name.split("/") if p]
    _STATE.scope.extend(parts)
    try:
        yield
    finally:
        del _STATE.scope[len(_STATE.scope) - len(parts):]


@contextlib.contextmanager
def auto_forward(flag=True):
    yield


def get_parameter_or_create(name, shape=None, initializer=None, need_grad=True, as_need_grad=None):
    full = name[1:] if name.startswith("/") else "/".join(_STATE.scope + [name])
    _STATE.asked.append(full)
    if full in _STATE.params:
        p = _STATE.params[full]
        if shape is not None:
            assert tuple(p.shape) == tuple(shape), (full, tuple(p.shape), tuple(shape))
        return p
    if _STATE.strict:
        raise KeyError(f"the reference asks for parameter {full!r} {tuple(shape)}: not in the registry "
                       f"(ndjir_b200/nnabla_names.py does not produce it)")
    if callable(initializer):
        val = np.asarray(initializer(tuple(shape)))
    else:
        val = np.asarray(initializer if initializer is not None else np.zeros(shape))
    p = V(val.reshape(tuple(shape)), requires_grad=bool(need_grad))
    _STATE.params[full] = p
    return p


def get_parameters(params=None, path="", grad_only=True):
    return OrderedDict(_STATE.params)


def grad(outputs, inputs, grad_outputs=None, **kw):
    gos = [torch.ones_like(o) for o in outputs] if grad_outputs is None else grad_outputs
    gs = torch.autograd.grad(list(outputs), list(inputs), grad_outputs=gos, create_graph=True, allow_unused=True)
    return [W_(g) for g in gs]


def set_parameters(named, trainable=lambda n: not n.endswith("photogrammetric-light-network/gain")):
    _STATE.params.clear()
    for k, v in named.items():
        _STATE.params[k] = V(v, requires_grad=trainable(k))
    _STATE.asked.clear()


def set_randoms(rand_by_seed, randn_by_seed):
    """{seed: [arrays]} - arrays are handed out by matching shape"""
    _STATE.randoms = {"rand": rand_by_seed, "randn": randn_by_seed}


def _draw(kind, shape, seed):
    cands = _STATE.randoms[kind].get(seed, [])
    for a in cands:
        if tuple(a.shape) == tuple(shape):
            return V(a)
    raise KeyError(f"F.{kind}(shape={tuple(shape)}, seed={seed}): no explicit random tensor registered")


# ----------------------------------------------------------------------------------------------------
# F.*
# ----------------------------------------------------------------------------------------------------
def _make_F():
    F = types.ModuleType("nnabla.functions")

    def idx_long(i):
        return i.detach().as_subclass(torch.Tensor).long() if i.is_floating_point() else i.as_subclass(torch.Tensor)

    def softplus(x, beta=1.0):
        bx = x * beta
        return W_((torch.clamp(bx, min=0) + torch.log1p(torch.exp(-torch.abs(bx)))) / beta)

    def cumprod(x, axis=None, exclusive=False, reverse=False):
        assert not reverse
        c = torch.cumprod(x, dim=axis)
        if exclusive:
            one = torch.ones_like(x.narrow(axis, 0, 1))
            c = torch.cat([one, c.narrow(axis, 0, x.shape[axis] - 1)], dim=axis)
        return W_(c)

    def gather(x, indices, axis=None, batch_dims=0):
        assert axis == batch_dims, "only the form used by sampler.py:225-234"
        idx = idx_long(indices)
        n = x.shape[axis]
        idx = torch.clamp(idx, 0, n - 1)                     # q14: out-of-range gather is undefined in nnabla
        extra = x.ndim - axis - 1
        ii = idx.reshape(idx.shape + (1,) * extra).expand(idx.shape + tuple(x.shape[axis + 1:]))
        return W_(torch.gather(x, axis, ii))

    def clip_by_value(x, min, max):
        lo = min if torch.is_tensor(min) else None
        hi = max if torch.is_tensor(max) else None
        y = x
        y = torch.maximum(y, lo.expand_as(y)) if lo is not None else torch.clamp(y, min=float(min))
        y = torch.minimum(y, hi.expand_as(y)) if hi is not None else torch.clamp(y, max=float(max))
        return W_(y)

    def _sum(x, axis=None, keepdims=False):
        if axis is None:
            return W_(x.sum())
        return W_(x.sum(dim=axis, keepdim=keepdims))

    def _mean(x, axis=None, keepdims=False):
        return W_(x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims))

    def concatenate(*xs, axis=None):
        return W_(torch.cat(list(xs), dim=(xs[0].ndim - 1 if axis is None else axis)))

    def bce(x, t):
        return W_(-(t * torch.log(x) + (1 - t) * torch.log(1 - x)))

    F.softplus = softplus
    F.cumprod = cumprod
    F.gather = gather
    F.clip_by_value = clip_by_value
    F.sum = _sum
    F.mean = _mean
    F.concatenate = concatenate
    F.binary_cross_entropy = bce
    F.cumsum = lambda x, axis=None, exclusive=False, reverse=False: W_(torch.cumsum(x, dim=axis))
    F.searchsorted = lambda seq, v, right=False: torch.searchsorted(seq.detach().contiguous(), v.detach().contiguous(),
                                                                    right=right)
    F.sort = lambda x, axis=-1, reverse=False, with_index=False, only_index=False: W_(torch.sort(x, dim=axis)[0])
    F.reshape = lambda x, shape, inplace=True: W_(x.reshape(tuple(shape)))
    F.broadcast = lambda x, shape: W_(x.expand(tuple(shape)))
    F.stack = lambda *xs, axis=0: W_(torch.stack(list(xs), dim=axis))
    F.min = lambda x, axis=None, keepdims=False, **kw: W_(torch.min(x, dim=axis, keepdim=keepdims)[0])
    F.norm = lambda x, p=None, axis=None, keepdims=False: W_(torch.sqrt((x * x).sum(dim=axis, keepdim=keepdims)))
    F.sigmoid = lambda x: W_(torch.sigmoid(x))
    F.relu = lambda x, inplace=False: W_(torch.relu(x))
    F.exp = lambda x: W_(torch.exp(x))
    F.log = lambda x: W_(torch.log(x))
    F.cos = lambda x: W_(torch.cos(x))
    F.sin = lambda x: W_(torch.sin(x))
    F.identity = lambda x: W_(x * 1.0)
    F.absolute_error = lambda a, b: W_(torch.abs(a - b))
    F.squared_error = lambda a, b: W_((a - b) ** 2)
    F.constant = lambda val=0, shape=(): torch.full(tuple(shape), float(val), dtype=DT).as_subclass(Variable)
    F.arange = lambda start, stop, step=1: torch.arange(start, stop, step, dtype=DT).as_subclass(Variable)
    F.greater_scalar = lambda x, val=1.0: W_((x > val).to(DT))
    F.maximum_scalar = lambda x, val=1.0: W_(torch.clamp(x, min=val))
    F.rand = lambda low=0, high=1, shape=(), seed=-1: _draw("rand", shape, seed)
    F.randn = lambda mu=0, sigma=1, shape=(), seed=-1: _draw("randn", shape, seed) * sigma + mu
    return F


def _make_PF(F):
    PF = types.ModuleType("nnabla.parametric_functions")

    def affine(inp, n_outmaps, base_axis=1, w_init=None, b_init=None, fix_parameters=False, rng=None, with_bias=True,
               apply_w=None, name=None):
        assert apply_w is None, "use_wn is false in every BASELINE config"
        assert base_axis == inp.ndim - 1
        with parameter_scope("affine"):
            W = get_parameter_or_create("W", (inp.shape[-1], n_outmaps), w_init, True, not fix_parameters)
            b = get_parameter_or_create("b", (n_outmaps,), b_init, True, not fix_parameters) if with_bias else None
        y = inp @ W
        return W_(y + b if b is not None else y)

    PF.affine = affine
    PF.weight_normalization = None
    return PF


# ----------------------------------------------------------------------------------------------------
# native boundary: grid queries, TV loss, ray bounds, direction sampling  -> oracle of the native kernels
# ----------------------------------------------------------------------------------------------------
def _install_native(F, PF):
    from oracle import cpu_ref as R
    from oracle import cpu_render as CR

    def pf_grid(scope, fn, shape_of):
        def q(x, G, feature_size, min_=[-1, -1, -1], max_=[1, 1, 1], use_ste=False, f_init=None, fix_parameters=False,
              rng=None):
            with parameter_scope(scope):
                Fp = get_parameter_or_create("F", shape_of(G, feature_size), f_init, True, not fix_parameters)
            # use_ste: the backward the wrapper registers for nn.grad returns (None, None) (voxel_feature.py:383-391,
            # triplane_feature.py / triline_feature.py alike): no gradient reaches the query
            return W_(fn(x.detach() if use_ste else x, Fp, 1.0))
        return q
    PF.query_on_voxel = pf_grid("voxel_feature", CR.voxel_query_torch,
                                lambda G, D: tuple([G] * 3 if isinstance(G, int) else G) + (D,))
    PF.query_on_triplane = pf_grid("triplane_feature", CR.triplane_query_torch, lambda G, D: (3, G, G, D))
    PF.query_on_triline = pf_grid("triline_feature", CR.triline_query_torch, lambda G, D: (3, G, D))
    for fam in ("cosine", "lanczos"):      # named in network.py:138-148's dispatch table; not on a BASELINE config
        for g in ("voxel", "triplane", "triline"):
            setattr(PF, f"{fam}_query_on_{g}", None)

    def tv(fn):
        def f(query, feature, min_=[-1, -1, -1], max_=[1, 1, 1], sym_backward=False, boundary_check=False, ctx=None):
            return W_(CR._tv_with_sym(fn, query.detach(), feature, sym_backward))
        return f
    F.tv_loss_on_voxel = tv(CR.tv_voxel_torch)
    F.tv_loss_on_triplane = tv(CR.tv_triplane_torch)
    F.tv_loss_on_triline = tv(CR.tv_triline_torch)
    F.tv_loss_on_voxel_hash = None

    def ray_aabb_intersection(camloc, raydir, min=[-1.0] * 3, max=[1.0] * 3, ctx=None):
        tn, tf, nh = R.ray_aabb(camloc.detach().numpy().astype(np.float32), raydir.detach().numpy().astype(np.float32),
                                list(min), list(max))
        B, Rr, _ = raydir.shape
        return tuple(V(np.asarray(a, np.float64).reshape(B, Rr, 1)) for a in (tn, tf, nh))

    def ray_sphere_intersection(camloc, raydir, radius, ctx=None):
        tn, tf, nh = R.ray_sphere(camloc.detach().numpy().astype(np.float32), raydir.detach().numpy().astype(np.float32),
                                  radius)
        B, Rr, _ = raydir.shape
        return tuple(V(np.asarray(a, np.float64).reshape(B, Rr, 1)) for a in (tn, tf, nh))

    def sample_uniform_directions(normal, cdf_the, cdf_phi, eps=0.0, ctx=None):
        return V(R.sample_directions(normal.detach().numpy().astype(np.float32), cdf_the.detach().numpy().astype(np.float32),
                                     cdf_phi.detach().numpy().astype(np.float32)))

    def sample_importance_directions(normal, cdf_the, cdf_phi, alpha, eps=0.0, ctx=None):
        return V(R.sample_directions(normal.detach().numpy().astype(np.float32), cdf_the.detach().numpy().astype(np.float32),
                                     cdf_phi.detach().numpy().astype(np.float32),
                                     alpha.detach().numpy().astype(np.float32)))
    return dict(ray_aabb_intersection=ray_aabb_intersection, ray_sphere_intersection=ray_sphere_intersection,
                sample_uniform_directions=sample_uniform_directions,
                sample_importance_directions=sample_importance_directions)


class _Adam:
    """nnabla.solvers.Adam as documented (see oracle/cpu_solver.py header); float64 here."""

    def __init__(self, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.b1, self.b2, self.eps, self.t = alpha, beta1, beta2, eps, 0
        self.params, self.state = OrderedDict(), {}

    def set_parameters(self, params, reset=True, retain_state=False):
        self.params = OrderedDict(params)

    def set_learning_rate(self, lr):
        self.alpha = lr

    def learning_rate(self):
        return self.alpha

    def zero_grad(self):
        for p in self.params.values():
            p.grad = torch.zeros_like(p.detach())

    def weight_decay(self, rate):
        for p in self.params.values():
            p.grad = p.grad + rate * p.detach()

    def clip_grad_by_norm(self, n):
        raise NotImplementedError

    def check_inf_or_nan_grad(self):
        return any(not bool(torch.isfinite(p.grad).all()) for p in self.params.values())

    def update(self):
        self.t += 1
        a_t = self.alpha * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for k, p in self.params.items():
                m, v = self.state.setdefault(k, (torch.zeros_like(p), torch.zeros_like(p)))
                m.mul_(self.b1).add_((1 - self.b1) * p.grad)
                v.mul_(self.b2).add_((1 - self.b2) * p.grad * p.grad)
                p.sub_(a_t * m / (torch.sqrt(v) + self.eps))


# ----------------------------------------------------------------------------------------------------
# module installation + import of the reference's files
# ----------------------------------------------------------------------------------------------------
_LOADED = {}


def load_reference_modules():
    """Returns {name: module} for sampler, network, renderer, specular_brdf, loss, solver imported from REF/python."""
    if _LOADED:
        return _LOADED
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} not found")
    nn = types.ModuleType("nnabla")
    F = _make_F()
    PF = _make_PF(F)
    nn.functions, nn.parametric_functions = F, PF
    nn.Variable = nn.NdArray = Variable
    nn.parameter_scope, nn.auto_forward = parameter_scope, auto_forward
    nn.get_parameters, nn.grad = get_parameters, grad
    nn.parameter = types.SimpleNamespace(get_parameter_or_create=get_parameter_or_create)
    nn.backward_functions = types.SimpleNamespace(register=lambda name, fn: None)
    nn.logger = types.SimpleNamespace(info=lambda *a, **k: None)
    nn.get_current_context = lambda: None
    init = types.ModuleType("nnabla.initializer")
    init.BaseInitializer = object
    init.ConstantInitializer = lambda v=0: (lambda shape: np.full(shape, v))
    init.NormalInitializer = lambda s=1.0, rng=None: (lambda shape: np.zeros(shape))
    rnd = types.ModuleType("nnabla.random")
    rnd.prng = np.random.RandomState(313)
    fn = types.ModuleType("nnabla.function")
    fn.PythonFunction = PythonFunction
    solvers = types.ModuleType("nnabla.solvers")
    solvers.Adam = _Adam
    nn.solvers = solvers
    mods = {"nnabla": nn, "nnabla.functions": F, "nnabla.parametric_functions": PF, "nnabla.initializer": init,
            "nnabla.random": rnd, "nnabla.function": fn, "nnabla.solvers": solvers,
            "inverse_transform_cuda": types.ModuleType("inverse_transform_cuda")}
    native = _install_native(F, PF)
    # packages whose modules only register native-kernel wrappers into F / PF (done above) or need cv2 / hydra
    gf = types.ModuleType("grid_feature")
    gf.__path__ = []
    mods["grid_feature"] = gf
    for m in ("cosine_triline_feature", "cosine_triplane_feature", "cosine_voxel_feature", "lanczos_triline_feature",
              "lanczos_triplane_feature", "lanczos_voxel_feature", "triline_feature", "triplane_feature",
              "voxel_feature", "total_variation_loss", "total_variation_loss_on_triline",
              "total_variation_loss_on_triplane", "total_variation_loss_on_voxel_hash"):
        mods[f"grid_feature.{m}"] = types.ModuleType(f"grid_feature.{m}")
        setattr(gf, m, mods[f"grid_feature.{m}"])
    inter = types.ModuleType("intersection")
    inter.__path__ = []
    ia = types.ModuleType("intersection.ray_aabb_intersection")
    ia.ray_aabb_intersection = native["ray_aabb_intersection"]
    isph = types.ModuleType("intersection.ray_sphere_intersection")
    isph.ray_sphere_intersection = native["ray_sphere_intersection"]
    mods.update({"intersection": inter, "intersection.ray_aabb_intersection": ia,
                 "intersection.ray_sphere_intersection": isph})
    helper = types.ModuleType("helper")          # helper.py imports cv2 / hydra; renderer.py only needs two names
    helper.generate_all_pixels = helper.generate_raydir_camloc = None
    mods["helper"] = helper
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    sys.path.insert(0, os.path.join(REF, "python"))
    try:
        for name in ("network", "sampler", "specular_brdf", "renderer", "loss", "solver"):
            assert name not in sys.modules or name in _LOADED, f"module name clash: {name}"
            _LOADED[name] = importlib.import_module(name)
            assert _LOADED[name].__file__.startswith(REF), _LOADED[name].__file__
    finally:
        sys.path.remove(os.path.join(REF, "python"))
    # light directions come from a CUDA kernel in the reference (sampler.py:376-387): route to the native-boundary oracle
    _LOADED["renderer"].sample_uniform_directions = native["sample_uniform_directions"]
    _LOADED["renderer"].sample_importance_directions = native["sample_importance_directions"]
    _LOADED["nn"], _LOADED["F"] = nn, F
    return _LOADED
