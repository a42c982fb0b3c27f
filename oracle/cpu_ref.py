"""CPU restatement (numpy, fp32 with the reference's operation order) of NDJIR's native kernels.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ndjir_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker / CPU baseline - never as the product path.

Every function cites the reference file:line it follows (paths relative to the reference
checkout).  Pinning status:
  * ray_aabb / ray_sphere / direction sampling: pinned against the numpy oracles the
    reference keeps inside its own tests (tests/golden/*.npz, made by tests/golden/make_golden.py).
  * grid families (voxel, triplane, triline, lanczos_voxel, TV): pinned against the reference's
    own ``*_composite.py`` statements executed through a numpy stand-in for ``nnabla.functions``
    (tests/golden/make_golden.py) and, on the GPU box, against the reference's own .cu files
    compiled unmodified into ``oracle/_ref`` (oracle/build_ref.py).
  * voxel_hash: the reference pins its hash only against itself; here it is pinned against
    ``oracle/_ref`` on the GPU box.

Value conventions: inputs are float32; "exact" quantities (cell indices, hash indices, hit
counts) are computed with IEEE fp32 arithmetic in the reference's order so they are
bit-identical to the CUDA kernels.  Scatter-adds accumulate in float64 (the CUDA atomics are
order-nondeterministic; float64 is the order-free answer they are compared to).
"""
import numpy as np

f32 = np.float32


def _f3(v):
    return np.asarray(v, dtype=f32).reshape(3)


# --------------------------------------------------------------------------------------
# shared cell arithmetic: csrc/grid_feature/voxel_feature_cuda.cu:52-70 (identical text in every
# linear family: triplane_feature_cuda.cu:56-67, triline_feature_cuda.cu:52-63, voxel_hash :146-162)
# --------------------------------------------------------------------------------------
_INTERP = ["linear"]


class interp:
    """`with interp("cosine"):` switches cell() to the weights of the cosine_* families
    (csrc/grid_feature/cosine_voxel_feature_cuda.cu:52-67, :163; cosine_triplane_feature_cuda.cu:65, :145;
    cosine_triline_feature_cuda.cu:61, :138): pqr0 = 0.5 cos(pi (xyz - xyz0)) + 0.5 and a per-point derivative scale
    scales * 0.5 pi sin(pi (xyz - xyz0)).  Everything else (cells, corner order, scatter signs) is shared."""

    def __init__(self, kind):
        assert kind in ("linear", "cosine")
        self.kind = kind

    def __enter__(self):
        _INTERP.append(self.kind)

    def __exit__(self, *exc):
        _INTERP.pop()


def _s(s, axis):
    """derivative scale on `axis` broadcastable against (B, D): a float (linear) or a (B,1) column (cosine)"""
    s = np.asarray(s, dtype=np.float64)
    return float(s[axis]) if s.ndim == 1 else s[:, axis][:, None]


def _s1(s, axis):
    """the same, broadcastable against (B,)"""
    s = np.asarray(s, dtype=np.float64)
    return float(s[axis]) if s.ndim == 1 else s[:, axis]


def cell(query, grid_sizes, min_, max_):
    """Returns (xyz0 uint32, xyz1 uint32, pqr0, pqr1, scales, xyz) for queries (B,3)."""
    q = np.asarray(query, dtype=f32).reshape(-1, 3)
    g1 = np.asarray(grid_sizes, dtype=f32).reshape(-1) - f32(1.0)
    if g1.size == 1:
        g1 = np.repeat(g1, 3)
    mn, mx = _f3(min_), _f3(max_)
    scales = (g1 / (mx - mn)).astype(f32)           # IEEE fp32 division
    xyz = ((q - mn).astype(f32) * scales).astype(f32)  # separate subtract, then multiply
    xyz0 = np.floor(xyz)
    xyz0 = np.maximum(xyz0, f32(0.0))
    xyz0 = np.minimum(xyz0, g1)
    xyz1 = np.minimum(xyz0 + f32(1.0), g1)
    if _INTERP[-1] == "cosine":
        ang = (f32(np.pi) * (xyz - xyz0).astype(f32)).astype(f32)
        pqr0 = (f32(0.5) * np.cos(ang).astype(f32) + f32(0.5)).astype(f32)
        pqr1 = (f32(1.0) - pqr0).astype(f32)
        dscale = (scales[None, :] * (f32(0.5 * np.pi) * np.sin(ang).astype(f32))).astype(f32)   # (B,3)
        return xyz0.astype(np.uint32), xyz1.astype(np.uint32), pqr0, pqr1, dscale, xyz
    pqr0 = (xyz1 - xyz).astype(f32)
    pqr1 = (f32(1.0) - pqr0).astype(f32)
    return xyz0.astype(np.uint32), xyz1.astype(np.uint32), pqr0, pqr1, scales, xyz


def _corners8(i0, i1, w0, w1):
    """Corner enumeration in the reference's order 000,001,010,011,100,101,110,111 (x,y,z bits)."""
    out = []
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                ix = (i1 if cx else i0)[:, 0]
                iy = (i1 if cy else i0)[:, 1]
                iz = (i1 if cz else i0)[:, 2]
                wx = (w1 if cx else w0)[:, 0]
                wy = (w1 if cy else w0)[:, 1]
                wz = (w1 if cz else w0)[:, 2]
                out.append((cx, cy, cz, ix, iy, iz, wx, wy, wz))
    return out


# --------------------------------------------------------------------------------------
# voxel family: csrc/grid_feature/voxel_feature_cuda.cu
# --------------------------------------------------------------------------------------
def voxel_indices(query, grid_sizes, D, min_, max_):
    """Flat feature indices of the 8 corners (channel 0), uint32, shape (B,8).
    voxel_feature_cuda.cu:78-80 (x*stride_x + y*stride_y + z*stride_z + d)."""
    G = np.asarray(grid_sizes, dtype=np.int64).reshape(3)
    i0, i1, p0, p1, s, _ = cell(query, G, min_, max_)
    sx, sy, sz = int(G[1] * G[2] * D), int(G[2] * D), int(D)
    idx = [c[3].astype(np.int64) * sx + c[4].astype(np.int64) * sy + c[5].astype(np.int64) * sz
           for c in _corners8(i0, i1, p0, p1)]
    return np.stack(idx, axis=1).astype(np.uint32)


def voxel_query(query, feature, min_, max_):
    """kernel_query_on_voxel, voxel_feature_cuda.cu:34-97.  feature (Gx,Gy,Gz,D) -> (B,D)."""
    feature = np.asarray(feature, dtype=f32)
    G = feature.shape[:3]
    i0, i1, p0, p1, s, _ = cell(query, G, min_, max_)
    out = np.zeros((i0.shape[0], feature.shape[3]), dtype=f32)
    for (_, _, _, ix, iy, iz, wx, wy, wz) in _corners8(i0, i1, p0, p1):
        w = ((wx * wy).astype(f32) * wz).astype(f32)
        out = (out + w[:, None] * feature[ix, iy, iz]).astype(f32)
    return out


def _voxel_corner_values(feature, i0, i1):
    f = {}
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                ix = (i1 if cx else i0)[:, 0]
                iy = (i1 if cy else i0)[:, 1]
                iz = (i1 if cz else i0)[:, 2]
                f[(cx, cy, cz)] = feature[ix, iy, iz].astype(np.float64)
    return f


def _voxel_dfdq(query, feature, min_, max_):
    """Per-channel spatial gradient (B,D,3) of the trilinear interpolant incl. the (G-1)/(max-min)
    scale, voxel_feature_cuda.cu:181-199."""
    feature = np.asarray(feature, dtype=f32)
    G = feature.shape[:3]
    i0, i1, p0, p1, s, _ = cell(query, G, min_, max_)
    f = _voxel_corner_values(feature, i0, i1)
    P0, P1 = p0.astype(np.float64), p1.astype(np.float64)
    a = lambda v: v[:, None]
    px0, py0, pz0 = a(P0[:, 0]), a(P0[:, 1]), a(P0[:, 2])
    px1, py1, pz1 = a(P1[:, 0]), a(P1[:, 1]), a(P1[:, 2])
    gx = _s(s, 0) * (py0 * pz0 * (f[1, 0, 0] - f[0, 0, 0]) + py0 * pz1 * (f[1, 0, 1] - f[0, 0, 1])
                 + py1 * pz0 * (f[1, 1, 0] - f[0, 1, 0]) + py1 * pz1 * (f[1, 1, 1] - f[0, 1, 1]))
    gy = _s(s, 1) * (px0 * pz0 * (f[0, 1, 0] - f[0, 0, 0]) + px0 * pz1 * (f[0, 1, 1] - f[0, 0, 1])
                 + px1 * pz0 * (f[1, 1, 0] - f[1, 0, 0]) + px1 * pz1 * (f[1, 1, 1] - f[1, 0, 1]))
    gz = _s(s, 2) * (px0 * py0 * (f[0, 0, 1] - f[0, 0, 0]) + px0 * py1 * (f[0, 1, 1] - f[0, 1, 0])
                 + px1 * py0 * (f[1, 0, 1] - f[1, 0, 0]) + px1 * py1 * (f[1, 1, 1] - f[1, 1, 0]))
    return np.stack([gx, gy, gz], axis=-1)


def voxel_grad_query(grad_output, query, feature, min_, max_):
    """kernel_grad_query, voxel_feature_cuda.cu:125-205 -> (B,3), summed over channels."""
    g = _voxel_dfdq(query, feature, min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(g.shape[0], g.shape[1])
    return (go[:, :, None] * g).sum(axis=1)


def voxel_grad_feature(grad_output, query, grid_sizes, D, min_, max_, out=None):
    """kernel_grad_feature, voxel_feature_cuda.cu:231-288 (scatter of ograd*p*q*r), float64 accumulate."""
    G = tuple(int(g) for g in grid_sizes)
    i0, i1, p0, p1, s, _ = cell(query, G, min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D)
    gf = np.zeros(G + (D,), dtype=np.float64) if out is None else out
    for (_, _, _, ix, iy, iz, wx, wy, wz) in _corners8(i0, i1, p0.astype(np.float64), p1.astype(np.float64)):
        np.add.at(gf, (ix, iy, iz), go * (wx * wy * wz)[:, None])
    return gf


def voxel_grad_query_grad_grad_output(grad_grad_query, query, feature, min_, max_):
    """kernel_grad_query_grad_grad_output, voxel_feature_cuda.cu:330-414: ggo_d = gg . d f_d/dq."""
    g = _voxel_dfdq(query, feature, min_, max_)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[:, None, :]).sum(axis=-1)


def voxel_grad_query_grad_query(grad_grad_query, grad_output, query, feature, min_, max_):
    """kernel_grad_query_grad_query, voxel_feature_cuda.cu:441-522 (cross second derivatives only)."""
    feature = np.asarray(feature, dtype=f32)
    G = feature.shape[:3]
    i0, i1, p0, p1, s, _ = cell(query, G, min_, max_)
    f = _voxel_corner_values(feature, i0, i1)
    P0, P1 = p0.astype(np.float64), p1.astype(np.float64)
    a = lambda v: v[:, None]
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], -1)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    sx, sy, sz = (float(v) for v in s)
    ti = go * sy * sz * (a(P0[:, 0]) * (f[0, 0, 0] - f[0, 0, 1] - f[0, 1, 0] + f[0, 1, 1])
                         + a(P1[:, 0]) * (f[1, 0, 0] - f[1, 0, 1] - f[1, 1, 0] + f[1, 1, 1]))
    tj = go * sx * sz * (a(P0[:, 1]) * (f[0, 0, 0] - f[0, 0, 1] - f[1, 0, 0] + f[1, 0, 1])
                         + a(P1[:, 1]) * (f[0, 1, 0] - f[0, 1, 1] - f[1, 1, 0] + f[1, 1, 1]))
    tk = go * sx * sy * (a(P0[:, 2]) * (f[0, 0, 0] - f[0, 1, 0] - f[1, 0, 0] + f[1, 1, 0])
                         + a(P1[:, 2]) * (f[0, 0, 1] - f[0, 1, 1] - f[1, 0, 1] + f[1, 1, 1]))
    ggx, ggy, ggz = a(gg[:, 0]), a(gg[:, 1]), a(gg[:, 2])
    gx = (ggy * tk + ggz * tj).sum(axis=1)
    gy = (ggz * ti + ggx * tk).sum(axis=1)
    gz = (ggx * tj + ggy * ti).sum(axis=1)
    return np.stack([gx, gy, gz], axis=-1)


def voxel_grad_query_grad_feature(grad_grad_query, grad_output, query, grid_sizes, D, min_, max_, out=None):
    """kernel_grad_query_grad_feature, voxel_feature_cuda.cu:549-614 (signed corner weights :603-611)."""
    G = tuple(int(g) for g in grid_sizes)
    i0, i1, p0, p1, s, _ = cell(query, G, min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros(G + (D,), dtype=np.float64) if out is None else out
    sx, sy, sz = _s1(s, 0), _s1(s, 1), _s1(s, 2)
    for (cx, cy, cz, ix, iy, iz, wx, wy, wz) in _corners8(i0, i1, p0.astype(np.float64), p1.astype(np.float64)):
        sgx, sgy, sgz = (1.0 if cx else -1.0), (1.0 if cy else -1.0), (1.0 if cz else -1.0)
        coef = gg[:, 0] * sx * (sgx * wy * wz) + gg[:, 1] * sy * (sgy * wx * wz) + gg[:, 2] * sz * (sgz * wx * wy)
        np.add.at(gf, (ix, iy, iz), go * coef[:, None])
    return gf


def voxel_grad_feature_grad_grad_output(grad_grad_feature, query, min_, max_):
    """kernel_grad_feature_grad_grad_output, voxel_feature_cuda.cu:642-710: interpolation of ggf."""
    return voxel_query(query, grad_grad_feature, min_, max_)


def voxel_grad_feature_grad_query(grad_grad_feature, grad_output, query, min_, max_):
    """kernel_grad_feature_grad_query, voxel_feature_cuda.cu:734-815: grad_query with ggf as the table."""
    return voxel_grad_query(grad_output, query, grad_grad_feature, min_, max_)


# --------------------------------------------------------------------------------------
# triplane / triline: csrc/grid_feature/triplane_feature_cuda.cu, triline_feature_cuda.cu,
# common_triplane.cuh:24-35 (plane 0=(x,y), 1=(y,z), 2=(z,x)), common.cuh:29-35 (n -> b,d,i)
# --------------------------------------------------------------------------------------
_PLANE_AXES = ((0, 1), (1, 2), (2, 0))


def triplane_query(query, feature, min_, max_):
    """kernel_query_on_triplane, triplane_feature_cuda.cu:38-90. feature (3,G,G,D) -> (B, D*3), c=d*3+i."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[3]
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    B = i0.shape[0]
    out = np.zeros((B, D, 3), dtype=f32)
    for i, (au, av) in enumerate(_PLANE_AXES):
        u0, u1, v0, v1 = i0[:, au], i1[:, au], i0[:, av], i1[:, av]
        a0, a1, b0, b1 = p0[:, au], p1[:, au], p0[:, av], p1[:, av]
        F = feature[i]
        f = ((a0 * b0).astype(f32)[:, None] * F[u0, v0]).astype(f32)
        f = (f + (a0 * b1).astype(f32)[:, None] * F[u0, v1]).astype(f32)
        f = (f + (a1 * b0).astype(f32)[:, None] * F[u1, v0]).astype(f32)
        f = (f + (a1 * b1).astype(f32)[:, None] * F[u1, v1]).astype(f32)
        out[:, :, i] = f
    return out.reshape(B, D * 3)


def _triplane_dfdq(query, feature, min_, max_):
    """(B,D,3planes,3axes) gradient, triplane_feature_cuda.cu:159-171."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[3]
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    B = i0.shape[0]
    g = np.zeros((B, D, 3, 3), dtype=np.float64)
    for i, (au, av) in enumerate(_PLANE_AXES):
        u0, u1, v0, v1 = i0[:, au], i1[:, au], i0[:, av], i1[:, av]
        a0, a1, b0, b1 = (p0[:, au].astype(np.float64)[:, None], p1[:, au].astype(np.float64)[:, None],
                          p0[:, av].astype(np.float64)[:, None], p1[:, av].astype(np.float64)[:, None])
        F = feature[i].astype(np.float64)
        f00, f01, f10, f11 = F[u0, v0], F[u0, v1], F[u1, v0], F[u1, v1]
        g[:, :, i, au] = _s(s, au) * (b0 * (f10 - f00) + b1 * (f11 - f01))
        g[:, :, i, av] = _s(s, av) * (a0 * (f01 - f00) + a1 * (f11 - f10))
    return g


def triplane_grad_query(grad_output, query, feature, min_, max_):
    """kernel_grad_query, triplane_feature_cuda.cu:115-172."""
    g = _triplane_dfdq(query, feature, min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(g.shape[0], g.shape[1], 3)
    return (go[..., None] * g).sum(axis=(1, 2))


def triplane_grad_feature(grad_output, query, G, D, min_, max_, out=None):
    """kernel_grad_feature, triplane_feature_cuda.cu:203-255."""
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D, 3)
    gf = np.zeros((3, G, G, D), dtype=np.float64) if out is None else out
    P0, P1 = p0.astype(np.float64), p1.astype(np.float64)
    for i, (au, av) in enumerate(_PLANE_AXES):
        for (u, a) in ((i0[:, au], P0[:, au]), (i1[:, au], P1[:, au])):
            for (v, b) in ((i0[:, av], P0[:, av]), (i1[:, av], P1[:, av])):
                np.add.at(gf[i], (u, v), go[:, :, i] * (a * b)[:, None])
    return gf


def triplane_grad_query_grad_grad_output(grad_grad_query, query, feature, min_, max_):
    """kernel_grad_query_grad_grad_output, triplane_feature_cuda.cu:298-366 -> (B, D*3)."""
    g = _triplane_dfdq(query, feature, min_, max_)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[:, None, None, :]).sum(axis=-1).reshape(g.shape[0], -1)


def triplane_grad_query_grad_feature(grad_grad_query, grad_output, query, G, D, min_, max_, out=None):
    """kernel_grad_query_grad_feature, triplane_feature_cuda.cu:497-560."""
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D, 3)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros((3, G, G, D), dtype=np.float64) if out is None else out
    P0, P1 = p0.astype(np.float64), p1.astype(np.float64)
    for i, (au, av) in enumerate(_PLANE_AXES):
        ggu, ggv, su, sv = gg[:, au], gg[:, av], _s1(s, au), _s1(s, av)
        for cu, (u, a) in enumerate(((i0[:, au], P0[:, au]), (i1[:, au], P1[:, au]))):
            for cv, (v, b) in enumerate(((i0[:, av], P0[:, av]), (i1[:, av], P1[:, av]))):
                coef = ggu * su * ((1.0 if cu else -1.0) * b) + ggv * sv * ((1.0 if cv else -1.0) * a)
                np.add.at(gf[i], (u, v), go[:, :, i] * coef[:, None])
    return gf


def triline_query(query, feature, min_, max_):
    """kernel_query_on_triline, triline_feature_cuda.cu:35-83. feature (3,G,D) -> (B, D*3), c=d*3+i."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[2]
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    B = i0.shape[0]
    out = np.zeros((B, D, 3), dtype=f32)
    for i in range(3):
        f = (p0[:, i][:, None] * feature[i][i0[:, i]]).astype(f32)
        f = (f + p1[:, i][:, None] * feature[i][i1[:, i]]).astype(f32)
        out[:, :, i] = f
    return out.reshape(B, D * 3)


def _triline_dfdq(query, feature, min_, max_):
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[2]
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    g = np.zeros((i0.shape[0], D, 3), dtype=np.float64)   # line i only moves along axis i
    for i in range(3):
        F = feature[i].astype(np.float64)
        g[:, :, i] = _s(s, i) * (F[i1[:, i]] - F[i0[:, i]])
    return g


def triline_grad_query(grad_output, query, feature, min_, max_):
    """kernel_grad_query, triline_feature_cuda.cu:109-160."""
    g = _triline_dfdq(query, feature, min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(g.shape)
    return (go * g).sum(axis=1)


def triline_grad_feature(grad_output, query, G, D, min_, max_, out=None):
    """kernel_grad_feature, triline_feature_cuda.cu:187-232."""
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D, 3)
    gf = np.zeros((3, G, D), dtype=np.float64) if out is None else out
    for i in range(3):
        np.add.at(gf[i], i0[:, i], go[:, :, i] * p0[:, i].astype(np.float64)[:, None])
        np.add.at(gf[i], i1[:, i], go[:, :, i] * p1[:, i].astype(np.float64)[:, None])
    return gf


def triline_grad_query_grad_grad_output(grad_grad_query, query, feature, min_, max_):
    """kernel_grad_query_grad_grad_output, triline_feature_cuda.cu:277-335."""
    g = _triline_dfdq(query, feature, min_, max_)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[:, None, :]).reshape(g.shape[0], -1)


def triline_grad_query_grad_feature(grad_grad_query, grad_output, query, G, D, min_, max_, out=None):
    """kernel_grad_query_grad_feature, triline_feature_cuda.cu:468-520: -/+ ograd*ggu*su at u0/u1."""
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D, 3)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros((3, G, D), dtype=np.float64) if out is None else out
    for i in range(3):
        v = go[:, :, i] * (gg[:, i] * _s1(s, i))[:, None]
        np.add.at(gf[i], i0[:, i], -v)
        np.add.at(gf[i], i1[:, i], v)
    return gf


# --------------------------------------------------------------------------------------
# Lanczos voxel: csrc/grid_feature/lanczos_voxel_feature_cuda.cu, common.cuh:54-97
# --------------------------------------------------------------------------------------
# cosine_voxel / cosine_triplane / cosine_triline: csrc/grid_feature/cosine_{voxel,triplane,triline}_feature_cuda.cu
# (5 exports each).  Same kernels as the linear families with the cosine cell (see `interp`).
# --------------------------------------------------------------------------------------
def _cosine(fn):
    def wrapped(*args, **kw):
        with interp("cosine"):
            return fn(*args, **kw)
    wrapped.__name__ = "cosine_" + fn.__name__
    wrapped.__doc__ = "cosine-weight variant of " + fn.__name__ + ": " + (fn.__doc__ or "")
    return wrapped


cosine_voxel_query = _cosine(voxel_query)
cosine_voxel_grad_query = _cosine(voxel_grad_query)
cosine_voxel_grad_feature = _cosine(voxel_grad_feature)
cosine_voxel_grad_query_grad_grad_output = _cosine(voxel_grad_query_grad_grad_output)
cosine_voxel_grad_query_grad_feature = _cosine(voxel_grad_query_grad_feature)
cosine_triplane_query = _cosine(triplane_query)
cosine_triplane_grad_query = _cosine(triplane_grad_query)
cosine_triplane_grad_feature = _cosine(triplane_grad_feature)
cosine_triplane_grad_query_grad_grad_output = _cosine(triplane_grad_query_grad_grad_output)
cosine_triplane_grad_query_grad_feature = _cosine(triplane_grad_query_grad_feature)
cosine_triline_query = _cosine(triline_query)
cosine_triline_grad_query = _cosine(triline_grad_query)
cosine_triline_grad_feature = _cosine(triline_grad_feature)
cosine_triline_grad_query_grad_grad_output = _cosine(triline_grad_query_grad_grad_output)
cosine_triline_grad_query_grad_feature = _cosine(triline_grad_query_grad_feature)


# --------------------------------------------------------------------------------------
def _sinc32(x):
    """common.cuh:54-59 (argument already narrowed to float)."""
    x = np.asarray(x, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        y = (np.sin(x) / x).astype(f32)
    return np.where(x == 0, f32(1.0), y).astype(f32)


def lanczos_weight(x, a=2):
    """common.cuh:62-69: z = M_PI * x is a double product, narrowed when passed to sinc(float)."""
    x = np.asarray(x, dtype=f32)
    z = np.pi * x.astype(np.float64)
    u = _sinc32(z.astype(f32))
    v = _sinc32((z / a).astype(f32))
    return (u * v).astype(f32)


def lanczos_grad_coefficient(x, a=2):
    """common.cuh:82-97."""
    x = np.asarray(x, dtype=f32)
    z0 = np.pi * x.astype(np.float64)
    z1 = np.pi * x.astype(np.float64) / a
    s0, s1 = _sinc32(z0.astype(f32)), _sinc32(z1.astype(f32))
    t0 = ((np.cos(z0.astype(f32)).astype(f32) - s0) * s1).astype(f32)
    t1 = ((np.cos(z1.astype(f32)).astype(f32) - s1) * s0).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        gc = ((t0 + t1).astype(f32) / x).astype(f32)
    return np.where(x == 0, f32(0.0), gc).astype(f32)


def _lanczos_taps(query, grid_sizes, min_, max_, w=2):
    """Tap coordinates (clamped, lanczos_voxel_feature_cuda.cu:67-76), weights and derivative weights,
    each (B,3,2w).  xyz0 = floor(xyz) is NOT clamped (:61)."""
    q = np.asarray(query, dtype=f32).reshape(-1, 3)
    g1 = np.asarray(grid_sizes, dtype=f32).reshape(3) - f32(1.0)
    mn, mx = _f3(min_), _f3(max_)
    scales = (g1 / (mx - mn)).astype(f32)
    xyz = ((q - mn).astype(f32) * scales).astype(f32)
    xyz0 = np.floor(xyz)
    offs = np.arange(-w + 1, w + 1, dtype=f32)
    taps = np.clip(xyz0[:, :, None] + offs[None, None, :], f32(0.0), g1[None, :, None]).astype(f32)
    d = (xyz[:, :, None] - taps).astype(f32)
    c = lanczos_weight(d, w)
    gc = lanczos_grad_coefficient(d, w)
    return taps.astype(np.uint32), c, gc, scales


def lanczos_voxel_indices(query, grid_sizes, D, min_, max_, w=2):
    """Flat indices (channel 0) of the (2w)^3 taps in loop order i,j,k; uint32 (B,64)."""
    G = np.asarray(grid_sizes, dtype=np.int64).reshape(3)
    taps, _, _, _ = _lanczos_taps(query, G, min_, max_, w)
    t = taps.astype(np.int64)
    idx = (t[:, 0, :, None, None] * (G[1] * G[2] * D) + t[:, 1, None, :, None] * (G[2] * D)
           + t[:, 2, None, None, :] * D)
    return idx.reshape(idx.shape[0], -1).astype(np.uint32)


def lanczos_voxel_query(query, feature, min_, max_, w=2):
    """kernel_query_on_voxel, lanczos_voxel_feature_cuda.cu:36-84."""
    feature = np.asarray(feature, dtype=f32)
    taps, c, _, _ = _lanczos_taps(query, feature.shape[:3], min_, max_, w)
    B, K = taps.shape[0], 2 * w
    out = np.zeros((B, feature.shape[3]), dtype=f32)
    for i in range(K):
        for j in range(K):
            for k in range(K):
                cijk = ((c[:, 0, i] * c[:, 1, j]).astype(f32) * c[:, 2, k]).astype(f32)
                out = (out + cijk[:, None] * feature[taps[:, 0, i], taps[:, 1, j], taps[:, 2, k]]).astype(f32)
    return out


def _lanczos_dfdq(query, feature, min_, max_, w=2):
    feature = np.asarray(feature, dtype=f32)
    taps, c, gc, s = _lanczos_taps(query, feature.shape[:3], min_, max_, w)
    B, K = taps.shape[0], 2 * w
    g = np.zeros((B, feature.shape[3], 3), dtype=np.float64)
    C, GC = c.astype(np.float64), gc.astype(np.float64)
    for i in range(K):
        for j in range(K):
            for k in range(K):
                f = feature[taps[:, 0, i], taps[:, 1, j], taps[:, 2, k]].astype(np.float64)
                g[:, :, 0] += (float(s[0]) * GC[:, 0, i] * C[:, 1, j] * C[:, 2, k])[:, None] * f
                g[:, :, 1] += (float(s[1]) * C[:, 0, i] * GC[:, 1, j] * C[:, 2, k])[:, None] * f
                g[:, :, 2] += (float(s[2]) * C[:, 0, i] * C[:, 1, j] * GC[:, 2, k])[:, None] * f
    return g


def lanczos_voxel_grad_query(grad_output, query, feature, min_, max_, w=2):
    """kernel_grad_query, lanczos_voxel_feature_cuda.cu:120-180."""
    g = _lanczos_dfdq(query, feature, min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(g.shape[0], g.shape[1])
    return (go[:, :, None] * g).sum(axis=1)


def lanczos_voxel_grad_feature(grad_output, query, grid_sizes, D, min_, max_, w=2, out=None):
    """kernel_grad_feature, lanczos_voxel_feature_cuda.cu:218-275."""
    G = tuple(int(g) for g in grid_sizes)
    taps, c, _, _ = _lanczos_taps(query, G, min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(taps.shape[0], D)
    gf = np.zeros(G + (D,), dtype=np.float64) if out is None else out
    C, K = c.astype(np.float64), 2 * w
    for i in range(K):
        for j in range(K):
            for k in range(K):
                np.add.at(gf, (taps[:, 0, i], taps[:, 1, j], taps[:, 2, k]),
                          go * (C[:, 0, i] * C[:, 1, j] * C[:, 2, k])[:, None])
    return gf


def lanczos_voxel_grad_query_grad_grad_output(grad_grad_query, query, feature, min_, max_, w=2):
    """kernel_grad_query_grad_grad_output, lanczos_voxel_feature_cuda.cu:319-377."""
    g = _lanczos_dfdq(query, feature, min_, max_, w)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[:, None, :]).sum(axis=-1)


def lanczos_voxel_grad_query_grad_feature(grad_grad_query, grad_output, query, grid_sizes, D, min_, max_, w=2,
                                          out=None):
    """kernel_grad_query_grad_feature, lanczos_voxel_feature_cuda.cu:523-585."""
    G = tuple(int(g) for g in grid_sizes)
    taps, c, gc, s = _lanczos_taps(query, G, min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(taps.shape[0], D)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros(G + (D,), dtype=np.float64) if out is None else out
    C, GC, K = c.astype(np.float64), gc.astype(np.float64), 2 * w
    for i in range(K):
        for j in range(K):
            for k in range(K):
                coef = (gg[:, 0] * float(s[0]) * GC[:, 0, i] * C[:, 1, j] * C[:, 2, k]
                        + gg[:, 1] * float(s[1]) * C[:, 0, i] * GC[:, 1, j] * C[:, 2, k]
                        + gg[:, 2] * float(s[2]) * C[:, 0, i] * C[:, 1, j] * GC[:, 2, k])
                np.add.at(gf, (taps[:, 0, i], taps[:, 1, j], taps[:, 2, k]), go * coef[:, None])
    return gf


# --------------------------------------------------------------------------------------
# lanczos triplane / triline: csrc/grid_feature/lanczos_triplane_feature_cuda.cu, lanczos_triline_feature_cuda.cu
# (4x4 / 4 taps per plane / line with the same clamped tap coordinates and window functions as the voxel family;
# output (B, D*3), c = d*3 + plane)
# --------------------------------------------------------------------------------------
def lanczos_triplane_query(query, feature, min_, max_, w=2):
    """kernel_query_on_triplane, lanczos_triplane_feature_cuda.cu:38-83."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[3]
    taps, c, _, _ = _lanczos_taps(query, (G, G, G), min_, max_, w)
    B, K = taps.shape[0], 2 * w
    out = np.zeros((B, D, 3), dtype=f32)
    for l, (au, av) in enumerate(_PLANE_AXES):
        f = np.zeros((B, D), dtype=f32)
        for i in range(K):
            for j in range(K):
                cij = (c[:, au, i] * c[:, av, j]).astype(f32)
                f = (f + cij[:, None] * feature[l][taps[:, au, i], taps[:, av, j]]).astype(f32)
        out[:, :, l] = f
    return out.reshape(B, D * 3)


def _lanczos_triplane_dfdq(query, feature, min_, max_, w=2):
    """(B, D, 3 planes, 3 axes), lanczos_triplane_feature_cuda.cu:149-162."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[3]
    taps, c, gc, s = _lanczos_taps(query, (G, G, G), min_, max_, w)
    B, K = taps.shape[0], 2 * w
    C, GC = c.astype(np.float64), gc.astype(np.float64)
    g = np.zeros((B, D, 3, 3), dtype=np.float64)
    for l, (au, av) in enumerate(_PLANE_AXES):
        for i in range(K):
            for j in range(K):
                f = feature[l][taps[:, au, i], taps[:, av, j]].astype(np.float64)
                g[:, :, l, au] += (float(s[au]) * GC[:, au, i] * C[:, av, j])[:, None] * f
                g[:, :, l, av] += (float(s[av]) * C[:, au, i] * GC[:, av, j])[:, None] * f
    return g


def lanczos_triplane_grad_query(grad_output, query, feature, min_, max_, w=2):
    """kernel_grad_query, lanczos_triplane_feature_cuda.cu:117-175."""
    g = _lanczos_triplane_dfdq(query, feature, min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(g.shape[0], g.shape[1], 3)
    return (go[..., None] * g).sum(axis=(1, 2))


def lanczos_triplane_grad_feature(grad_output, query, G, D, min_, max_, w=2, out=None):
    """kernel_grad_feature, lanczos_triplane_feature_cuda.cu:208-262."""
    taps, c, _, _ = _lanczos_taps(query, (G, G, G), min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(taps.shape[0], D, 3)
    gf = np.zeros((3, G, G, D), dtype=np.float64) if out is None else out
    C, K = c.astype(np.float64), 2 * w
    for l, (au, av) in enumerate(_PLANE_AXES):
        for i in range(K):
            for j in range(K):
                np.add.at(gf[l], (taps[:, au, i], taps[:, av, j]), go[:, :, l] * (C[:, au, i] * C[:, av, j])[:, None])
    return gf


def lanczos_triplane_grad_query_grad_grad_output(grad_grad_query, query, feature, min_, max_, w=2):
    """kernel_grad_query_grad_grad_output, lanczos_triplane_feature_cuda.cu:304-365 -> (B, D*3)."""
    g = _lanczos_triplane_dfdq(query, feature, min_, max_, w)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[:, None, None, :]).sum(axis=-1).reshape(g.shape[0], -1)


def lanczos_triplane_grad_query_grad_feature(grad_grad_query, grad_output, query, G, D, min_, max_, w=2, out=None):
    """kernel_grad_query_grad_feature, lanczos_triplane_feature_cuda.cu:503-560."""
    taps, c, gc, s = _lanczos_taps(query, (G, G, G), min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(taps.shape[0], D, 3)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros((3, G, G, D), dtype=np.float64) if out is None else out
    C, GC, K = c.astype(np.float64), gc.astype(np.float64), 2 * w
    for l, (au, av) in enumerate(_PLANE_AXES):
        for i in range(K):
            for j in range(K):
                coef = (gg[:, au] * float(s[au]) * GC[:, au, i] * C[:, av, j]
                        + gg[:, av] * float(s[av]) * C[:, au, i] * GC[:, av, j])
                np.add.at(gf[l], (taps[:, au, i], taps[:, av, j]), go[:, :, l] * coef[:, None])
    return gf


def lanczos_triline_query(query, feature, min_, max_, w=2):
    """kernel_query_on_triline, lanczos_triline_feature_cuda.cu:37-77."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[2]
    taps, c, _, _ = _lanczos_taps(query, (G, G, G), min_, max_, w)
    B, K = taps.shape[0], 2 * w
    out = np.zeros((B, D, 3), dtype=f32)
    for l in range(3):
        f = np.zeros((B, D), dtype=f32)
        for i in range(K):
            f = (f + c[:, l, i][:, None] * feature[l][taps[:, l, i]]).astype(f32)
        out[:, :, l] = f
    return out.reshape(B, D * 3)


def _lanczos_triline_dfdq(query, feature, min_, max_, w=2):
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[2]
    taps, c, gc, s = _lanczos_taps(query, (G, G, G), min_, max_, w)
    g = np.zeros((taps.shape[0], D, 3), dtype=np.float64)      # line l only moves along axis l
    for l in range(3):
        for i in range(2 * w):
            g[:, :, l] += (float(s[l]) * gc[:, l, i].astype(np.float64))[:, None] * feature[l][taps[:, l, i]]
    return g


def lanczos_triline_grad_query(grad_output, query, feature, min_, max_, w=2):
    """kernel_grad_query, lanczos_triline_feature_cuda.cu:107-150."""
    g = _lanczos_triline_dfdq(query, feature, min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(g.shape)
    return (go * g).sum(axis=1)


def lanczos_triline_grad_feature(grad_output, query, G, D, min_, max_, w=2, out=None):
    """kernel_grad_feature, lanczos_triline_feature_cuda.cu:182-228."""
    taps, c, _, _ = _lanczos_taps(query, (G, G, G), min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(taps.shape[0], D, 3)
    gf = np.zeros((3, G, D), dtype=np.float64) if out is None else out
    for l in range(3):
        for i in range(2 * w):
            np.add.at(gf[l], taps[:, l, i], go[:, :, l] * c[:, l, i].astype(np.float64)[:, None])
    return gf


def lanczos_triline_grad_query_grad_grad_output(grad_grad_query, query, feature, min_, max_, w=2):
    """kernel_grad_query_grad_grad_output, lanczos_triline_feature_cuda.cu:267-317."""
    g = _lanczos_triline_dfdq(query, feature, min_, max_, w)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[:, None, :]).reshape(g.shape[0], -1)


def lanczos_triline_grad_query_grad_feature(grad_grad_query, grad_output, query, G, D, min_, max_, w=2, out=None):
    """kernel_grad_query_grad_feature, lanczos_triline_feature_cuda.cu:455-505."""
    taps, c, gc, s = _lanczos_taps(query, (G, G, G), min_, max_, w)
    go = np.asarray(grad_output, dtype=np.float64).reshape(taps.shape[0], D, 3)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros((3, G, D), dtype=np.float64) if out is None else out
    for l in range(3):
        for i in range(2 * w):
            coef = gg[:, l] * float(s[l]) * gc[:, l, i].astype(np.float64)
            np.add.at(gf[l], taps[:, l, i], go[:, :, l] * coef[:, None])
    return gf


# --------------------------------------------------------------------------------------
# voxel hash: csrc/grid_feature/voxel_hash_feature_cuda.cu, common_voxel_hash.cuh:24-55
# --------------------------------------------------------------------------------------
def hash_force_align(size, mod=8):
    """common_voxel_hash.cuh:24-28 (sic: size + size % mod)."""
    return size + size % mod


def hash_level_table(G0, growth_factor, T0, L, D, Gs_override=None):
    """(G_l, T_l, offset_l) per level + total; common_voxel_hash.cuh:31-55 evaluated in fp32 like the
    device code (pow(float,int) -> float, float product, floor).  `Gs_override`: the level grid sizes as the
    DEVICE evaluated them (its pow is not exactly rounded, q2); tests pass what the GPU reports."""
    Gs, Ts, offs = [], [], []
    off = 0
    gf = f32(growth_factor)
    for l in range(L):
        p = f32(1.0)
        for _ in range(l):      # exact for growth factors whose powers are fp32-exact (1.5, 2.0)
            p = f32(p * gf)
        G = int(np.floor(f32(f32(G0) * p))) if Gs_override is None else int(Gs_override[l])
        Gf = f32(G)
        T = min(int(min(f32(f32(Gf * Gf) * Gf), f32(T0))), int(T0))
        Gs.append(G); Ts.append(T); offs.append(off)
        off += hash_force_align(T * D)
    return Gs, Ts, offs, off


def hash_index3(x, y, z, T):
    """voxel_hash_feature_cuda.cu:38-48 (tiny-cuda-nn primes), uint32 wrap-around then % T."""
    x = np.asarray(x, dtype=np.uint64); y = np.asarray(y, dtype=np.uint64); z = np.asarray(z, dtype=np.uint64)
    m = np.uint64(0xFFFFFFFF)
    r = (x & m) ^ ((y * np.uint64(2654435761)) & m) ^ ((z * np.uint64(805459861)) & m)
    return (r % np.uint64(T)).astype(np.uint32)


def hash_corner_indices(query, G, T, min_, max_):
    """kernel_hash_index, voxel_hash_feature_cuda.cu:55-101 -> (B,8) uint32 in corner order 000..111."""
    i0, i1, p0, p1, s, _ = cell(query, (G, G, G), min_, max_)
    return np.stack([hash_index3(c[3], c[4], c[5], T) for c in _corners8(i0, i1, p0, p1)], axis=1)


def voxel_hash_query(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None):
    """kernel_voxel_hash_feature, voxel_hash_feature_cuda.cu:124-195.  Returns the reference's
    (D,L,B) layout (:190); transpose(2,0,1).reshape(B,D*L) gives the wrapper's output (c=d*L+l)."""
    feature = np.asarray(feature, dtype=f32).reshape(-1)
    Gs, Ts, offs, _ = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    B = np.asarray(query).reshape(-1, 3).shape[0]
    out = np.zeros((D, L, B), dtype=f32)
    for l in range(L):
        i0, i1, p0, p1, s, _ = cell(query, (Gs[l],) * 3, min_, max_)
        tab = feature[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D)
        acc = np.zeros((B, D), dtype=f32)
        for (_, _, _, ix, iy, iz, wx, wy, wz) in _corners8(i0, i1, p0, p1):
            w = ((wx * wy).astype(f32) * wz).astype(f32)
            acc = (acc + w[:, None] * tab[hash_index3(ix, iy, iz, Ts[l])]).astype(f32)
        out[:, l, :] = acc.T
    return out


def _hash_dfdq(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None):
    """(D,L,B,3) spatial gradient, voxel_hash_feature_cuda.cu:273-293."""
    feature = np.asarray(feature, dtype=f32).reshape(-1)
    Gs, Ts, offs, _ = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    B = np.asarray(query).reshape(-1, 3).shape[0]
    g = np.zeros((D, L, B, 3), dtype=np.float64)
    for l in range(L):
        i0, i1, p0, p1, s, _ = cell(query, (Gs[l],) * 3, min_, max_)
        tab = feature[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D).astype(np.float64)
        f = {}
        for (cx, cy, cz, ix, iy, iz, _, _, _) in _corners8(i0, i1, p0, p1):
            f[(cx, cy, cz)] = tab[hash_index3(ix, iy, iz, Ts[l])]
        P0, P1 = p0.astype(np.float64), p1.astype(np.float64)
        a = lambda v: v[:, None]
        px0, py0, pz0 = a(P0[:, 0]), a(P0[:, 1]), a(P0[:, 2])
        px1, py1, pz1 = a(P1[:, 0]), a(P1[:, 1]), a(P1[:, 2])
        gx = _s(s, 0) * (py0 * pz0 * (f[1, 0, 0] - f[0, 0, 0]) + py0 * pz1 * (f[1, 0, 1] - f[0, 0, 1])
                     + py1 * pz0 * (f[1, 1, 0] - f[0, 1, 0]) + py1 * pz1 * (f[1, 1, 1] - f[0, 1, 1]))
        gy = _s(s, 1) * (px0 * pz0 * (f[0, 1, 0] - f[0, 0, 0]) + px0 * pz1 * (f[0, 1, 1] - f[0, 0, 1])
                     + px1 * pz0 * (f[1, 1, 0] - f[1, 0, 0]) + px1 * pz1 * (f[1, 1, 1] - f[1, 0, 1]))
        gz = _s(s, 2) * (px0 * py0 * (f[0, 0, 1] - f[0, 0, 0]) + px0 * py1 * (f[0, 1, 1] - f[0, 1, 0])
                     + px1 * py0 * (f[1, 0, 1] - f[1, 0, 0]) + px1 * py1 * (f[1, 1, 1] - f[1, 1, 0]))
        g[:, l, :, 0], g[:, l, :, 1], g[:, l, :, 2] = gx.T, gy.T, gz.T
    return g


def voxel_hash_grad_query(grad_output_dlb, query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None):
    """kernel_grad_query, voxel_hash_feature_cuda.cu:221-300; grad_output in (D,L,B) layout."""
    g = _hash_dfdq(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs)
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    return (go[..., None] * g).sum(axis=(0, 1))


def voxel_hash_grad_feature(grad_output_dlb, query, G0, growth_factor, T0, L, D, min_, max_, out=None, Gs=None):
    """kernel_grad_feature, voxel_hash_feature_cuda.cu:336-400."""
    Gs, Ts, offs, total = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    gf = np.zeros(total, dtype=np.float64) if out is None else out
    for l in range(L):
        i0, i1, p0, p1, s, _ = cell(query, (Gs[l],) * 3, min_, max_)
        tab = gf[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D)
        for (_, _, _, ix, iy, iz, wx, wy, wz) in _corners8(i0, i1, p0.astype(np.float64), p1.astype(np.float64)):
            np.add.at(tab, hash_index3(ix, iy, iz, Ts[l]), go[:, l, :].T * (wx * wy * wz)[:, None])
    return gf


def voxel_hash_grad_query_grad_grad_output(grad_grad_query, query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None):
    """kernel_grad_query_grad_grad_output, voxel_hash_feature_cuda.cu:446-540 -> (D,L,B)."""
    g = _hash_dfdq(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[None, None, :, :]).sum(axis=-1)


def voxel_hash_grad_query_grad_feature(grad_grad_query, grad_output_dlb, query, G0, growth_factor, T0, L, D,
                                       min_, max_, out=None, Gs=None):
    """kernel_grad_query_grad_feature, voxel_hash_feature_cuda.cu:673-750."""
    Gs, Ts, offs, total = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros(total, dtype=np.float64) if out is None else out
    for l in range(L):
        i0, i1, p0, p1, s, _ = cell(query, (Gs[l],) * 3, min_, max_)
        tab = gf[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D)
        sx, sy, sz = (float(v) for v in s)
        for (cx, cy, cz, ix, iy, iz, wx, wy, wz) in _corners8(i0, i1, p0.astype(np.float64), p1.astype(np.float64)):
            sgx, sgy, sgz = (1.0 if cx else -1.0), (1.0 if cy else -1.0), (1.0 if cz else -1.0)
            coef = (gg[:, 0] * sx * (sgx * wy * wz) + gg[:, 1] * sy * (sgy * wx * wz)
                    + gg[:, 2] * sz * (sgz * wx * wy))
            np.add.at(tab, hash_index3(ix, iy, iz, Ts[l]), go[:, l, :].T * coef[:, None])
    return gf


# --------------------------------------------------------------------------------------
# lanczos voxel hash: csrc/grid_feature/lanczos_voxel_hash_feature_cuda.cu (level table, offsets and (D,L,B)
# layout of the trilinear hash family; per level the 4x4x4 clamped Lanczos taps, each hashed)
# --------------------------------------------------------------------------------------
def lanczos_hash_cell_index(cells, T):
    """kernel_hash_index, lanczos_voxel_hash_feature_cuda.cu:54-66: one integer cell (as floats) per row."""
    c = np.asarray(cells, dtype=f32).reshape(-1, 3).astype(np.uint32)
    return hash_index3(c[:, 0], c[:, 1], c[:, 2], T)


def _lz_hash_levels(query, G0, growth_factor, T0, L, D, min_, max_, Gs, w=2):
    Gs, Ts, offs, total = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    for l in range(L):
        taps, c, gc, s = _lanczos_taps(query, (Gs[l],) * 3, min_, max_, w)
        yield l, Ts[l], offs[l], taps, c.astype(np.float64), gc.astype(np.float64), [float(v) for v in s], c
    return


def lanczos_voxel_hash_query(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None, w=2):
    """kernel_voxel_hash_feature, lanczos_voxel_hash_feature_cuda.cu:131-195 -> (D,L,B)."""
    feature = np.asarray(feature, dtype=f32).reshape(-1)
    B, K = np.asarray(query).reshape(-1, 3).shape[0], 2 * w
    out = np.zeros((D, L, B), dtype=f32)
    for l, T, off, taps, _, _, _, c32 in _lz_hash_levels(query, G0, growth_factor, T0, L, D, min_, max_, Gs, w):
        tab = feature[off: off + T * D].reshape(T, D)
        acc = np.zeros((B, D), dtype=f32)
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    cijk = ((c32[:, 0, i] * c32[:, 1, j]).astype(f32) * c32[:, 2, k]).astype(f32)
                    acc = (acc + cijk[:, None] * tab[hash_index3(taps[:, 0, i], taps[:, 1, j], taps[:, 2, k], T)]).astype(f32)
        out[:, l, :] = acc.T
    return out


def _lz_hash_dfdq(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None, w=2):
    """(D,L,B,3), lanczos_voxel_hash_feature_cuda.cu:262-287."""
    feature = np.asarray(feature, dtype=f32).reshape(-1)
    B, K = np.asarray(query).reshape(-1, 3).shape[0], 2 * w
    g = np.zeros((D, L, B, 3), dtype=np.float64)
    for l, T, off, taps, C, GC, s, _ in _lz_hash_levels(query, G0, growth_factor, T0, L, D, min_, max_, Gs, w):
        tab = feature[off: off + T * D].reshape(T, D).astype(np.float64)
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    f = tab[hash_index3(taps[:, 0, i], taps[:, 1, j], taps[:, 2, k], T)]      # (B,D)
                    g[:, l, :, 0] += ((s[0] * GC[:, 0, i] * C[:, 1, j] * C[:, 2, k])[:, None] * f).T
                    g[:, l, :, 1] += ((s[1] * C[:, 0, i] * GC[:, 1, j] * C[:, 2, k])[:, None] * f).T
                    g[:, l, :, 2] += ((s[2] * C[:, 0, i] * C[:, 1, j] * GC[:, 2, k])[:, None] * f).T
    return g


def lanczos_voxel_hash_grad_query(grad_output_dlb, query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None):
    """kernel_grad_query, lanczos_voxel_hash_feature_cuda.cu:222-300."""
    g = _lz_hash_dfdq(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs)
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    return (go[..., None] * g).sum(axis=(0, 1))


def lanczos_voxel_hash_grad_feature(grad_output_dlb, query, G0, growth_factor, T0, L, D, min_, max_, out=None, Gs=None,
                                    w=2):
    """kernel_grad_feature, lanczos_voxel_hash_feature_cuda.cu:327-385."""
    total = hash_level_table(G0, growth_factor, T0, L, D, Gs)[3]
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    gf = np.zeros(total, dtype=np.float64) if out is None else out
    K = 2 * w
    for l, T, off, taps, C, _, _, _ in _lz_hash_levels(query, G0, growth_factor, T0, L, D, min_, max_, Gs, w):
        tab = gf[off: off + T * D].reshape(T, D)
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    np.add.at(tab, hash_index3(taps[:, 0, i], taps[:, 1, j], taps[:, 2, k], T),
                              go[:, l, :].T * (C[:, 0, i] * C[:, 1, j] * C[:, 2, k])[:, None])
    return gf


def lanczos_voxel_hash_grad_query_grad_grad_output(grad_grad_query, query, feature, G0, growth_factor, T0, L, D, min_,
                                                   max_, Gs=None):
    """kernel_grad_query_grad_grad_output, lanczos_voxel_hash_feature_cuda.cu:437-512 -> (D,L,B)."""
    g = _lz_hash_dfdq(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    return (g * gg[None, None, :, :]).sum(axis=-1)


def lanczos_voxel_hash_grad_query_grad_feature(grad_grad_query, grad_output_dlb, query, G0, growth_factor, T0, L, D,
                                               min_, max_, out=None, Gs=None, w=2):
    """kernel_grad_query_grad_feature, lanczos_voxel_hash_feature_cuda.cu:649-705."""
    total = hash_level_table(G0, growth_factor, T0, L, D, Gs)[3]
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    gg = np.asarray(grad_grad_query, dtype=np.float64).reshape(-1, 3)
    gf = np.zeros(total, dtype=np.float64) if out is None else out
    K = 2 * w
    for l, T, off, taps, C, GC, s, _ in _lz_hash_levels(query, G0, growth_factor, T0, L, D, min_, max_, Gs, w):
        tab = gf[off: off + T * D].reshape(T, D)
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    coef = (gg[:, 0] * s[0] * GC[:, 0, i] * C[:, 1, j] * C[:, 2, k]
                            + gg[:, 1] * s[1] * C[:, 0, i] * GC[:, 1, j] * C[:, 2, k]
                            + gg[:, 2] * s[2] * C[:, 0, i] * C[:, 1, j] * GC[:, 2, k])
                    np.add.at(tab, hash_index3(taps[:, 0, i], taps[:, 1, j], taps[:, 2, k], T),
                              go[:, l, :].T * coef[:, None])
    return gf


# --------------------------------------------------------------------------------------
# total variation: csrc/grid_feature/total_variation_loss{,_on_triplane,_on_triline}_cuda.cu
# --------------------------------------------------------------------------------------
def tv_voxel(query, feature, min_, max_):
    """kernel_tv_loss_on_voxel_forward, total_variation_loss_cuda.cu:33-85 -> (B,D)."""
    feature = np.asarray(feature, dtype=f32)
    i0, i1, _, _, _, _ = cell(query, feature.shape[:3], min_, max_)
    f000 = feature[i0[:, 0], i0[:, 1], i0[:, 2]]
    dx = feature[i1[:, 0], i0[:, 1], i0[:, 2]] - f000
    dy = feature[i0[:, 0], i1[:, 1], i0[:, 2]] - f000
    dz = feature[i0[:, 0], i0[:, 1], i1[:, 2]] - f000
    return np.sqrt(((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32) + (dz * dz).astype(f32)).astype(f32)


def tv_voxel_backward(grad_output, query, feature, min_, max_, sym_backward, out=None):
    """kernel_tv_loss_on_voxel_backward, total_variation_loss_cuda.cu:111-174 (double intermediate :161)."""
    feature = np.asarray(feature, dtype=f32)
    i0, i1, _, _, _, _ = cell(query, feature.shape[:3], min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], -1)
    f000 = feature[i0[:, 0], i0[:, 1], i0[:, 2]]
    dx = (feature[i1[:, 0], i0[:, 1], i0[:, 2]] - f000).astype(f32)
    dy = (feature[i0[:, 0], i1[:, 1], i0[:, 2]] - f000).astype(f32)
    dz = (feature[i0[:, 0], i0[:, 1], i1[:, 2]] - f000).astype(f32)
    ss = ((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32) + (dz * dz).astype(f32)
    common = go / np.sqrt(ss.astype(np.float64) + 1e-12)
    gf = np.zeros(feature.shape, dtype=np.float64) if out is None else out
    g100, g010, g001 = common * dx, common * dy, common * dz
    np.add.at(gf, (i1[:, 0], i0[:, 1], i0[:, 2]), g100)
    np.add.at(gf, (i0[:, 0], i1[:, 1], i0[:, 2]), g010)
    np.add.at(gf, (i0[:, 0], i0[:, 1], i1[:, 2]), g001)
    if sym_backward:
        np.add.at(gf, (i0[:, 0], i0[:, 1], i0[:, 2]), -(g100 + g010 + g001))
    return gf


def tv_voxel_hash(query, feature, G0, growth_factor, T0, L, D, min_, max_, Gs=None):
    """kernel_tv_loss_on_voxel_hash_forward, total_variation_loss_on_voxel_hash_cuda.cu:50-106 -> (D,L,B)."""
    feature = np.asarray(feature, dtype=f32).reshape(-1)
    Gs, Ts, offs, _ = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    B = np.asarray(query).reshape(-1, 3).shape[0]
    out = np.zeros((D, L, B), dtype=f32)
    for l in range(L):
        i0, i1, _, _, _, _ = cell(query, (Gs[l],) * 3, min_, max_)
        tab = feature[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D)
        f000 = tab[hash_index3(i0[:, 0], i0[:, 1], i0[:, 2], Ts[l])]
        dx = tab[hash_index3(i1[:, 0], i0[:, 1], i0[:, 2], Ts[l])] - f000
        dy = tab[hash_index3(i0[:, 0], i1[:, 1], i0[:, 2], Ts[l])] - f000
        dz = tab[hash_index3(i0[:, 0], i0[:, 1], i1[:, 2], Ts[l])] - f000
        out[:, l, :] = np.sqrt((dx * dx + dy * dy + dz * dz).astype(f32)).astype(f32).T
    return out


def tv_voxel_hash_backward(grad_output_dlb, query, feature, G0, growth_factor, T0, L, D, min_, max_, out=None, Gs=None):
    """kernel_tv_loss_on_voxel_hash_backward, :132-196: ograd * delta * rsqrt(sum + 1e-12) to the three upper neighbours;
    the lower corner gets nothing (no sym_backward in this family)."""
    feature = np.asarray(feature, dtype=f32).reshape(-1)
    Gs, Ts, offs, total = hash_level_table(G0, growth_factor, T0, L, D, Gs)
    go = np.asarray(grad_output_dlb, dtype=np.float64).reshape(D, L, -1)
    gf = np.zeros(total, dtype=np.float64) if out is None else out
    for l in range(L):
        i0, i1, _, _, _, _ = cell(query, (Gs[l],) * 3, min_, max_)
        tab = feature[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D)
        gtab = gf[offs[l]: offs[l] + Ts[l] * D].reshape(Ts[l], D)
        h000 = hash_index3(i0[:, 0], i0[:, 1], i0[:, 2], Ts[l])
        hx = hash_index3(i1[:, 0], i0[:, 1], i0[:, 2], Ts[l])
        hy = hash_index3(i0[:, 0], i1[:, 1], i0[:, 2], Ts[l])
        hz = hash_index3(i0[:, 0], i0[:, 1], i1[:, 2], Ts[l])
        f000 = tab[h000]
        dx, dy, dz = tab[hx] - f000, tab[hy] - f000, tab[hz] - f000
        s2 = (dx * dx + dy * dy + dz * dz).astype(f32).astype(np.float64)
        common = go[:, l, :].T / np.sqrt(s2 + 1e-12)
        np.add.at(gtab, hx, common * dx)
        np.add.at(gtab, hy, common * dy)
        np.add.at(gtab, hz, common * dz)
    return gf


def tv_triplane(query, feature, min_, max_):
    """kernel_tv_loss_on_triplane_forward, total_variation_loss_on_triplane_cuda.cu:31-75 -> (B, D*3)."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[3]
    i0, i1, _, _, _, _ = cell(query, (G, G, G), min_, max_)
    out = np.zeros((i0.shape[0], D, 3), dtype=f32)
    for i, (au, av) in enumerate(_PLANE_AXES):
        F = feature[i]
        f00 = F[i0[:, au], i0[:, av]]
        du = (F[i1[:, au], i0[:, av]] - f00).astype(f32)
        dv = (F[i0[:, au], i1[:, av]] - f00).astype(f32)
        out[:, :, i] = np.sqrt((du * du).astype(f32) + (dv * dv).astype(f32))
    return out.reshape(i0.shape[0], D * 3)


def tv_triplane_backward(grad_output, query, feature, min_, max_, sym_backward, out=None):
    """kernel_tv_loss_on_triplane_backward, total_variation_loss_on_triplane_cuda.cu:101-160."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[3]
    i0, i1, _, _, _, _ = cell(query, (G, G, G), min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D, 3)
    gf = np.zeros(feature.shape, dtype=np.float64) if out is None else out
    for i, (au, av) in enumerate(_PLANE_AXES):
        F = feature[i]
        f00 = F[i0[:, au], i0[:, av]]
        du = (F[i1[:, au], i0[:, av]] - f00).astype(f32)
        dv = (F[i0[:, au], i1[:, av]] - f00).astype(f32)
        ss = (du * du).astype(f32) + (dv * dv).astype(f32)
        common = go[:, :, i] / np.sqrt(ss.astype(np.float64) + 1e-12)
        g10, g01 = common * du, common * dv
        np.add.at(gf[i], (i1[:, au], i0[:, av]), g10)
        np.add.at(gf[i], (i0[:, au], i1[:, av]), g01)
        if sym_backward:
            np.add.at(gf[i], (i0[:, au], i0[:, av]), -(g10 + g01))
    return gf


def tv_triline(query, feature, min_, max_):
    """kernel_tv_loss_on_triline_forward, total_variation_loss_on_triline_cuda.cu:31-72: sqrt(delta^2)."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[2]
    i0, i1, _, _, _, _ = cell(query, (G, G, G), min_, max_)
    out = np.zeros((i0.shape[0], D, 3), dtype=f32)
    for i in range(3):
        du = (feature[i][i1[:, i]] - feature[i][i0[:, i]]).astype(f32)
        out[:, :, i] = np.sqrt((du * du).astype(f32))
    return out.reshape(i0.shape[0], D * 3)


def tv_triline_backward(grad_output, query, feature, min_, max_, sym_backward, out=None):
    """kernel_tv_loss_on_triline_backward, total_variation_loss_on_triline_cuda.cu:98-152."""
    feature = np.asarray(feature, dtype=f32)
    G, D = feature.shape[1], feature.shape[2]
    i0, i1, _, _, _, _ = cell(query, (G, G, G), min_, max_)
    go = np.asarray(grad_output, dtype=np.float64).reshape(i0.shape[0], D, 3)
    gf = np.zeros(feature.shape, dtype=np.float64) if out is None else out
    for i in range(3):
        du = (feature[i][i1[:, i]] - feature[i][i0[:, i]]).astype(f32)
        ss = (du * du).astype(f32)
        g1 = go[:, :, i] / np.sqrt(ss.astype(np.float64) + 1e-12) * du
        np.add.at(gf[i], i1[:, i], g1)
        if sym_backward:
            np.add.at(gf[i], i0[:, i], -g1)
    return gf


# --------------------------------------------------------------------------------------
# ray bounds: csrc/intersection/ray_aabb_intersection_cuda.cu, ray_sphere_intersection_cuda.cu
# --------------------------------------------------------------------------------------
def _fma32(a, b, c):
    """fp32 fused multiply-add (nvcc contracts camloc + t*raydir, ray_aabb_intersection_cuda.cu:53-58)."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64)
            + np.asarray(c, dtype=np.float64)).astype(f32)


def ray_aabb(camloc, raydir, min_, max_):
    """kernel_ray_aabb_intersection, ray_aabb_intersection_cuda.cu:71-142.
    camloc (B,3), raydir (B,R,3) -> t_near, t_far, n_hits each (B,R,1) float32."""
    camloc = np.asarray(camloc, dtype=f32); raydir = np.asarray(raydir, dtype=f32)
    B, R, _ = raydir.shape
    o = np.broadcast_to(camloc.reshape(B, 1, 3), (B, R, 3)).reshape(-1, 3)
    d = raydir.reshape(-1, 3)
    mn, mx = _f3(min_), _f3(max_)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = (f32(1.0) / d).astype(f32)
        t = np.concatenate([((mx - o).astype(f32) * inv).astype(f32),
                            ((mn - o).astype(f32) * inv).astype(f32)], axis=1)      # (N,6): +x,+y,+z,-x,-y,-z
        n = o.shape[0]
        n_hits = np.zeros(n, dtype=np.int32)
        first = np.zeros(n, dtype=np.int64)
        last = np.zeros(n, dtype=np.int64)
        for i in range(6):
            ti = t[:, i]
            x = _fma32(ti[:, None], d, o)
            x[:, i % 3] = mx[i % 3] if i < 3 else mn[i % 3]        # snap (:60-66)
            ok = ~np.isinf(ti)
            ok &= (ti >= 0)
            ok &= np.all((x >= mn) & (x <= mx), axis=1)
            first = np.where(ok & (n_hits == 0), i, first)
            last = np.where(ok & (n_hits > 0), i, last)
            n_hits = n_hits + ok.astype(np.int32)
    ar = np.arange(n)
    ta, tb = t[ar, first], t[ar, last]
    t_near = np.where(n_hits >= 2, np.where(ta <= tb, ta, tb), f32(0.0)).astype(f32)
    t_far = np.where(n_hits >= 2, np.where(ta <= tb, tb, ta), np.where(n_hits == 1, ta, f32(0.0))).astype(f32)
    shp = (B, R, 1)
    return t_near.reshape(shp), t_far.reshape(shp), n_hits.astype(f32).reshape(shp)


def ray_sphere(camloc, raydir, radius):
    """kernel_ray_sphere_intersection, ray_sphere_intersection_cuda.cu:27-78 (fp32, dots in x,y,z order)."""
    camloc = np.asarray(camloc, dtype=f32); raydir = np.asarray(raydir, dtype=f32)
    B, R, _ = raydir.shape
    o = np.broadcast_to(camloc.reshape(B, 1, 3), (B, R, 3)).reshape(-1, 3)
    d = raydir.reshape(-1, 3)

    def dot(a, b):   # helper_math.h dot(): a.x*b.x + a.y*b.y + a.z*b.z; nvcc fuses the FIRST product of each sum:
        # fma(a.z, b.z, fma(a.x, b.x, a.y*b.y)) (SASS of the sm_100a build: FMUL, FFMA, FFMA)
        return _fma32(a[:, 2], b[:, 2], _fma32(a[:, 0], b[:, 0], (a[:, 1] * b[:, 1]).astype(f32)))
    r2 = f32(f32(radius) * f32(radius))
    cv, vv, cc = dot(o, d), dot(d, d), dot(o, o)
    X = (-cv).astype(f32)
    ccr = _fma32(np.full_like(cc, -f32(radius)), np.full_like(cc, f32(radius)), cc)    # cc - r*r as fma(-r, r, cc)
    Y = _fma32(cv, cv, -(vv * ccr).astype(f32))
    Zi = (f32(1.0) / vv).astype(f32)
    with np.errstate(invalid="ignore"):
        Ys = np.sqrt(np.maximum(Y, 0)).astype(f32)
    tn = ((X - Ys).astype(f32) * Zi).astype(f32)
    tf = ((X + Ys).astype(f32) * Zi).astype(f32)
    pos = tn >= 0
    t_near = np.where(Y > 0, np.where(pos, tn, f32(0.0)), np.where(Y == 0, (X * Zi).astype(f32), f32(0.0)))
    t_far = np.where(Y > 0, tf, np.where(Y == 0, (X * Zi).astype(f32), f32(0.0)))
    n_hits = np.where(Y > 0, np.where(pos, 2, 1), np.where(Y == 0, 1, 0))
    shp = (B, R, 1)
    return t_near.astype(f32).reshape(shp), t_far.astype(f32).reshape(shp), n_hits.astype(f32).reshape(shp)


# --------------------------------------------------------------------------------------
# light-direction sampling: csrc/sampling/inverse_transform_cuda.cu
# --------------------------------------------------------------------------------------
def sample_directions(normal, cdf_the, cdf_phi, alpha=None, eps=0.0):
    """kernel_sample_uniform_directions (:31-69) / kernel_sample_importance_directions (:94-136).
    normal (B,R,3), cdf_the (B,R,nt), cdf_phi (B,R,np)[, alpha (B,R,1)] -> (B,R,nt*np,3)."""
    normal = np.asarray(normal, dtype=f32)
    B, R, _ = normal.shape
    n = normal.reshape(-1, 3)
    ct = np.asarray(cdf_the, dtype=f32).reshape(B * R, -1)
    cp = np.asarray(cdf_phi, dtype=f32).reshape(B * R, -1)
    nt, nph = ct.shape[1], cp.shape[1]
    the = np.repeat(ct, nph, axis=1)                 # m_the = m / n_phis
    phi_u = np.tile(cp, (1, nt))                     # m_phi = m % n_phis
    phi = (2.0 * np.pi * phi_u.astype(np.float64)).astype(f32)     # double product narrowed (:48)
    if alpha is None:
        cos_t = the
    else:
        a = np.asarray(alpha, dtype=f32).reshape(B * R, 1)
        a2 = (a * a).astype(f32)
        cos_t = np.sqrt(((f32(1.0) - the).astype(f32)
                         / (((a2 - f32(1.0)).astype(f32) * the).astype(f32) + f32(1.0)).astype(f32)).astype(f32))
    sin_t = np.sqrt((f32(1.0) - (cos_t * cos_t).astype(f32)).astype(f32)).astype(f32)
    x = (sin_t * np.cos(phi).astype(f32)).astype(f32)
    y = (sin_t * np.sin(phi).astype(f32)).astype(f32)
    z = cos_t.astype(f32)
    nb = (n + f32(eps)).astype(np.float64)
    zax = nb / np.linalg.norm(nb, axis=1, keepdims=True)
    xr = np.stack([-nb[:, 1], nb[:, 0], np.zeros_like(nb[:, 0])], axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        xax = xr / np.linalg.norm(xr, axis=1, keepdims=True)
    yax = np.cross(zax, xax)
    out = (x[..., None].astype(np.float64) * xax[:, None, :] + y[..., None].astype(np.float64) * yax[:, None, :]
           + z[..., None].astype(np.float64) * zax[:, None, :])
    return out.astype(f32).reshape(B, R, nt * nph, 3)


# --------------------------------------------------------------------------------------
# squareplus: csrc/activation/squareplus_cuda.cu:30-60
# --------------------------------------------------------------------------------------
def squareplus_forward(x, b):
    x = np.asarray(x, dtype=f32)
    return (f32(0.5) * (x + np.sqrt((x * x).astype(f32) + f32(b)).astype(f32))).astype(f32)


def squareplus_backward(dy, x, b):
    x = np.asarray(x, dtype=f32); dy = np.asarray(dy, dtype=f32)
    return (dy * f32(0.5) * (f32(1.0) + x / np.sqrt((x * x).astype(f32) + f32(b)))).astype(f32)
