"""Operator-level API of the grid-feature queries, mirroring the reference's `F.query_on_*` / `F.tv_loss_on_*`
functions (python/grid_feature/voxel_feature.py:138-142, triplane_feature.py, triline_feature.py,
lanczos_voxel_feature.py, voxel_hash_feature.py:211-241, total_variation_loss*.py) on torch CUDA tensors.

The reference implements each op as an nnabla `PythonFunction` (forward_impl / backward_impl calling the native
module) plus a `*GradQuery` function registered as its double-backward (voxel_feature.py:170-263, :383-399).  Here the
same pair is a pair of `torch.autograd.Function`s whose forward / backward bodies are single C-ABI calls into
libndjir_b200.so, so `torch.autograd.grad(out, query, create_graph=True)` plays the role of `nn.grad` and a second
backward reaches grad_query_grad_grad_output / grad_query_grad_query / grad_query_grad_feature exactly like the
reference's registered backward.  torch supplies tensors, streams and the autograd tape only.

Signatures keep the reference's argument names: `query (..., 3)`, `feature` in the reference layout, `min_`, `max_`
lists of 3 floats, `use_ste`, `boundary_check` (accepted and ignored, as in the reference kernels).
"""
import torch

from ._lib import call


def _st():
    return torch.cuda.current_stream().cuda_stream


def _flat(query):
    q = query.reshape(-1, 3)
    return q if q.is_contiguous() else q.contiguous()


class _Family:
    """C-ABI names and the grid-spec arguments of one family."""

    def __init__(self, prefix, query_name, spec, out_channels, has_gq_gq=False):
        self.p, self.qn, self.spec, self.out_channels, self.has_gq_gq = prefix, query_name, spec, out_channels, has_gq_gq


def _voxel_spec(feature):
    return [list(feature.shape[:3]), feature.shape[3]]


_FAMILIES = {
    "voxel": _Family("ndjir_voxel_", "query_on_voxel", _voxel_spec, lambda f: f.shape[3], has_gq_gq=True),
    "lanczos_voxel": _Family("ndjir_lanczos_voxel_", "query_on_voxel", _voxel_spec, lambda f: f.shape[3]),
    "triplane": _Family("ndjir_triplane_", "query_on_triplane", lambda f: [f.shape[1], f.shape[3]],
                        lambda f: 3 * f.shape[3]),
    "triline": _Family("ndjir_triline_", "query_on_triline", lambda f: [f.shape[1], f.shape[2]],
                       lambda f: 3 * f.shape[2]),
}


class _Query(torch.autograd.Function):
    """QueryOnVoxel & co. (voxel_feature.py:27-135)."""

    @staticmethod
    def forward(ctx, query, feature, fam, min_, max_, use_ste):
        F = _FAMILIES[fam]
        q = _flat(query)
        B, C = q.shape[0], F.out_channels(feature)
        out = torch.empty((B, C), dtype=torch.float32, device=query.device)
        call(F.p + F.qn, B, out, q, feature, *F.spec(feature), list(min_), list(max_), 0, _st())
        ctx.save_for_backward(query, feature)
        ctx.fam, ctx.mm, ctx.use_ste = fam, (list(min_), list(max_)), use_ste
        return out.reshape(query.shape[:-1] + (C,))

    @staticmethod
    def backward(ctx, grad_output):
        query, feature = ctx.saved_tensors
        F = _FAMILIES[ctx.fam]
        go = grad_output.reshape(-1, grad_output.shape[-1]).contiguous()
        gq = gf = None
        if ctx.needs_input_grad[0] and not ctx.use_ste:      # use_ste: no gradient to the query (voxel_feature.py:390-391)
            gq = _GradQuery.apply(go, query, feature, ctx.fam, *ctx.mm).reshape(query.shape)
        if ctx.needs_input_grad[1]:
            gf = torch.zeros_like(feature)
            spec = F.spec(feature)
            call(F.p + "grad_feature", go.shape[0], gf, go, _flat(query), *spec, *ctx.mm, 1, _st())
        return gq, gf, None, None, None, None


class _GradQuery(torch.autograd.Function):
    """QueryOnVoxelGradQuery (voxel_feature.py:170-263): forward = grad_query, backward = the three second-order
    kernels."""

    @staticmethod
    def forward(ctx, grad_output, query, feature, fam, min_, max_):
        F = _FAMILIES[fam]
        q = _flat(query)
        out = torch.empty((q.shape[0], 3), dtype=torch.float32, device=query.device)
        call(F.p + "grad_query", q.shape[0], out, grad_output, q, feature, *F.spec(feature), min_, max_, 0, _st())
        ctx.save_for_backward(grad_output, query, feature)
        ctx.fam, ctx.mm = fam, (min_, max_)
        return out

    @staticmethod
    def backward(ctx, ggq):
        grad_output, query, feature = ctx.saved_tensors
        F = _FAMILIES[ctx.fam]
        q = _flat(query)
        ggq = ggq.contiguous()
        B = q.shape[0]
        spec = F.spec(feature)
        ggo = gq = gf = None
        if ctx.needs_input_grad[0]:
            ggo = torch.empty_like(grad_output)
            call(F.p + "grad_query_grad_grad_output", B, ggo, ggq, q, feature, *spec, *ctx.mm, 0, _st())
        if ctx.needs_input_grad[1] and F.has_gq_gq:
            gq = torch.zeros((B, 3), dtype=torch.float32, device=query.device)
            call(F.p + "grad_query_grad_query", B, gq, ggq, grad_output, q, feature, *spec, *ctx.mm, _st())
            gq = gq.reshape(query.shape)
        if ctx.needs_input_grad[2]:
            gf = torch.zeros_like(feature)
            call(F.p + "grad_query_grad_feature", B, gf, ggq, grad_output, q, *spec, *ctx.mm, _st())
        return ggo, gq, gf, None, None, None


def _check(query, feature):
    if not (query.is_cuda and feature.is_cuda):
        raise ValueError("ndjir_b200 grid features need CUDA tensors (there is no CPU path)")
    if query.dtype != torch.float32 or feature.dtype != torch.float32:
        raise ValueError("fp32 only, like the reference (data_ptr(np.float32, ctx))")
    if query.shape[-1] != 3:
        raise ValueError("query must be (..., 3)")      # voxel_feature.py:62-64
    return feature if feature.is_contiguous() else feature.contiguous()


def query_on_voxel(query, feature, min_, max_, use_ste=False, boundary_check=False, ctx=None):
    """F.query_on_voxel (voxel_feature.py:138-142): query (...,3), feature (Gx,Gy,Gz,D) -> (...,D)."""
    return _Query.apply(query, _check(query, feature), "voxel", min_, max_, use_ste)


def lanczos_query_on_voxel(query, feature, min_, max_, use_ste=False, boundary_check=False, ctx=None):
    """F.lanczos_query_on_voxel (lanczos_voxel_feature.py): 64-tap separable Lanczos-2."""
    return _Query.apply(query, _check(query, feature), "lanczos_voxel", min_, max_, use_ste)


def query_on_triplane(query, feature, min_, max_, use_ste=False, boundary_check=False, ctx=None):
    """F.query_on_triplane: feature (3,G,G,D) -> (..., D*3), channel c = d*3 + plane."""
    return _Query.apply(query, _check(query, feature), "triplane", min_, max_, use_ste)


def query_on_triline(query, feature, min_, max_, use_ste=False, boundary_check=False, ctx=None):
    """F.query_on_triline: feature (3,G,D) -> (..., D*3)."""
    return _Query.apply(query, _check(query, feature), "triline", min_, max_, use_ste)


class _TV(torch.autograd.Function):
    """tv_loss_on_{voxel,triplane,triline} (total_variation_loss*.py): forward sqrt(sum delta^2) per sample cell,
    backward scatters into the feature gradient; no gradient to the query."""

    @staticmethod
    def forward(ctx, query, feature, fam, min_, max_, sym_backward):
        F = _FAMILIES[fam]
        q = _flat(query)
        C = F.out_channels(feature)
        out = torch.empty((q.shape[0], C), dtype=torch.float32, device=query.device)
        call(f"ndjir_tv_loss_on_{fam}", q.shape[0], out, q, feature, *F.spec(feature), list(min_), list(max_), _st())
        ctx.save_for_backward(query, feature)
        ctx.fam, ctx.mm, ctx.sym = fam, (list(min_), list(max_)), bool(sym_backward)
        return out.reshape(query.shape[:-1] + (C,))

    @staticmethod
    def backward(ctx, grad_output):
        query, feature = ctx.saved_tensors
        F = _FAMILIES[ctx.fam]
        go = grad_output.reshape(-1, grad_output.shape[-1]).contiguous()
        gf = torch.zeros_like(feature)
        call(f"ndjir_tv_loss_on_{ctx.fam}_backward", go.shape[0], gf, go, _flat(query), feature, *F.spec(feature), *ctx.mm,
             int(ctx.sym), _st())
        return None, gf, None, None, None, None


def tv_loss_on_voxel(query, feature, min_, max_, sym_backward=True, boundary_check=False, ctx=None):
    return _TV.apply(query, _check(query, feature), "voxel", min_, max_, sym_backward)


def tv_loss_on_triplane(query, feature, min_, max_, sym_backward=True, boundary_check=False, ctx=None):
    return _TV.apply(query, _check(query, feature), "triplane", min_, max_, sym_backward)


def tv_loss_on_triline(query, feature, min_, max_, sym_backward=True, boundary_check=False, ctx=None):
    return _TV.apply(query, _check(query, feature), "triline", min_, max_, sym_backward)


class _Hash(torch.autograd.Function):
    """QueryOnVoxelHash (voxel_hash_feature.py:67-209), output (B, D*L) with channel c = d*L + l (what the reference
    produces after its in-place transpose, :152-155)."""

    @staticmethod
    def forward(ctx, query, feature, G0, growth_factor, T0, L, D, min_, max_):
        q = _flat(query)
        out = torch.empty((q.shape[0], D * L), dtype=torch.float32, device=query.device)
        ctx.spec = (int(G0), float(growth_factor), int(T0), int(L), int(D), list(min_), list(max_))
        call("ndjir_voxel_hash_voxel_hash_feature", q.shape[0], out, q, feature, *ctx.spec, 1, 0, _st())
        ctx.save_for_backward(query, feature)
        return out.reshape(query.shape[:-1] + (D * L,))

    @staticmethod
    def backward(ctx, grad_output):
        query, feature = ctx.saved_tensors
        go = grad_output.reshape(-1, grad_output.shape[-1]).contiguous()
        q = _flat(query)
        gq = gf = None
        if ctx.needs_input_grad[0]:
            gq = torch.empty((q.shape[0], 3), dtype=torch.float32, device=query.device)
            call("ndjir_voxel_hash_grad_query", q.shape[0], gq, go, q, feature, *ctx.spec, 1, 0, _st())
            gq = gq.reshape(query.shape)
        if ctx.needs_input_grad[1]:
            gf = torch.zeros_like(feature)
            call("ndjir_voxel_hash_grad_feature", q.shape[0], gf, go, q, *ctx.spec, 1, 1, _st())
        return gq, gf, None, None, None, None, None, None, None


def query_on_voxel_hash(query, feature, G0=16, growth_factor=1.5, T0=2 ** 15, L=16, D=2, min_=(-1, -1, -1),
                        max_=(1, 1, 1), boundary_check=False, ctx=None):
    """F.query_on_voxel_hash (voxel_hash_feature.py:211-241); feature is the flat (n_params,) table."""
    n = call("ndjir_voxel_hash_num_params", int(G0), float(growth_factor), int(T0), int(L), int(D))
    if feature.numel() != n:
        raise ValueError(f"feature must have {n} elements for this level table, got {feature.numel()}")
    return _Hash.apply(query, _check(query, feature), G0, growth_factor, T0, L, D, min_, max_)
