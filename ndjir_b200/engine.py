"""Host side of the fused per-ray path: parameter store, MLP passes and the train step, built ONLY from calls into
libndjir_b200.so (hand-written sm_100a kernels behind the C ABI).  torch is used for device memory and streams.

What it mirrors in the reference (same names / argument meaning):
  sample_points      python/sampler.py:311-314 (SamplePoints.forward_impl :256-299)
  geometric_network  python/network.py:154-232
  pb_render          python/renderer.py:32-209
  total_loss         python/loss.py:27-192   (forward AND backward: loss.forward() + loss.backward(),
                                              python/train.py:135-140)
The reference gets its backward from nnabla's autodiff (including the double-backward through nn.grad,
renderer.py:52); here the backward is hand-derived and every product runs in ndjir_gemm.

There is no CPU or PyTorch fallback: every stage raises NdjirError if the CUDA library is missing.
"""
import math

import numpy as np
import torch

from . import _lib
from .config import grid_channels
from .parallel import allgather_rows, allreduce_gradients, allreduce_mask_sum
from .scene import network_dims, pe_dim, NET_ORDER

EPI_BIAS, EPI_SOFTPLUS, EPI_ACCUM, EPI_MUL_S, EPI_ADJ, EPI_ATOMIC = range(6)
LOSS_NAMES = ["loss", "loss_rgb", "loss_eikonal", "loss_tv", "loss_mask", "prior_base_color", "prior_roughness",
              "prior_specular_reflectance", "reg_std_roughness", "reg_std_specular_reflectance"]
N_LOSSES = len(LOSS_NAMES)


def r4(n):
    return (n + 3) // 4 * 4


class Layer:
    """One affine layer in INTERNAL layout: W (K, ldw) row-major with ldw = round4(N); `rowmap[i]` is the internal
    row of reference input row i (the engine keeps head inputs as [feature | x | normal | ...] so the geometric
    feature block is 16-byte aligned and shared between heads; the reference order is [x | ... | feature | normal])."""

    def __init__(self, K, N, rowmap=None, col0=0, Nref=None):
        self.K, self.N = K, N
        self.ldw = r4(N)
        self.rowmap = np.arange(K) if rowmap is None else np.asarray(rowmap)
        self.col0 = col0            # first reference output column held by this layer (split last layers)
        self.w_off = self.b_off = self.t_off = -1
        self.ldt = r4(K)            # row stride of the transposed copy W^T (N, ldt)


class ParamStore:
    """Flat fp32 parameter and gradient buffers (one allocation each: the gradient buffer is what a data-parallel
    run all-reduces) with per-layer views in internal layout, plus import/export in the reference's layout
    (nnabla PF.affine: W (in,out), b (out); names as oracle/cpu_render.Model.parameters())."""

    def __init__(self, conf, device):
        self.conf, self.device = conf, device
        dims = network_dims(conf)
        g = conf.geometric_network
        Df = g.feature_size
        self.Df = Df
        self.nets = {}
        # geometric network: reference row order kept; last layer split into sdf column / feature columns
        geo = [Layer(di, do) for (di, do) in dims["geo"][:-1]]
        Kl = dims["geo"][-1][0]
        geo += [Layer(Kl, 1, col0=0), Layer(Kl, Df, col0=1)]
        self.nets["geo"] = geo

        def first_map(ref_blocks, internal_order):
            """ref_blocks: [(name, width)] in reference order; internal_order: names in internal order."""
            off, pos = {}, 0
            for name in internal_order:
                w = dict(ref_blocks)[name]
                off[name] = pos
                pos += w
            rm = []
            for name, w in ref_blocks:
                rm += list(range(off[name], off[name] + w))
            return np.asarray(rm), pos

        def head(name, ref_blocks, internal_order):
            d = dims[name]
            rm, K = first_map(ref_blocks, internal_order)
            assert K == d[0][0], (name, K, d[0][0])
            return [Layer(K, d[0][1], rowmap=rm)] + [Layer(di, do) for (di, do) in d[1:]]
        el, sv, pl, bg = (conf.environment_light_network, conf.soft_visibility_light_network,
                          conf.photogrammetric_light_network, conf.background_network)
        self.nets["bc"] = head("bc", [("x", 3), ("f", Df)], ["f", "x"])
        for n in ("ii", "ro", "sp"):
            self.nets[n] = head(n, [("x", 3), ("f", Df), ("n", 3)], ["f", "x", "n"])
        self.nets["pl"] = head("pl", [("x", 3), ("pe", pe_dim(3, pl.pe_bands)), ("f", Df), ("n", 3), ("inv", 1)],
                               ["f", "x", "n", "pe", "inv"])
        self.nets["sv"] = head("sv", [("x", 3), ("pe", pe_dim(3, sv.pe_bands)), ("f", Df), ("n", 3)],
                               ["f", "x", "n", "pe"])
        self.nets["el"] = [Layer(di, do) for (di, do) in dims["el"]]
        bg0 = [Layer(di, do) for (di, do) in dims["bg0"][:-1]]
        Kb = dims["bg0"][-1][0]
        bg0 += [Layer(Kb, 1, col0=0), Layer(Kb, bg.feature_size0, col0=1)]
        self.nets["bg0"] = bg0
        self.nets["bg1"] = head("bg1", [("x", 4), ("f", bg.feature_size0), ("v", 3), ("pe", pe_dim(3, bg.pe_bands1))],
                                ["f", "x", "v", "pe"])
        n = 0
        for name in NET_ORDER:
            for L in self.nets[name]:
                L.w_off = n; n += L.K * L.ldw
                L.b_off = n; n += r4(L.N)
        self.gain_off = n; n += 4
        self.n_mlp = n
        nt = 0
        for name in NET_ORDER:
            for L in self.nets[name]:
                L.t_off = nt; nt += r4(L.N) * L.ldt
        # The skip layer's input block (rows [n_prev, K) of its W: the 43 network inputs re-entering at layer 4) starts at
        # row 213, which is not 16-byte aligned inside the transposed copy: it gets a transposed copy of its own so that
        # its input-gradient product also sees a TMA-friendly MN-major operand (K-major fallback: 0.36 ms, this: 0.16 ms)
        self.skip_t = {}
        sk = list(g.skip_layers)[:1]
        for name in ("geo",):
            if sk and 0 < sk[0] < len(self.nets[name]):
                L, Lp = self.nets[name][sk[0]], self.nets[name][sk[0] - 1]
                rows = L.K - Lp.N
                if rows > 0 and Lp.N % 4 != 0:
                    self.skip_t[id(L)] = (Lp.N, rows, nt, r4(rows))
                    nt += r4(L.N) * r4(rows)
        # transposed weight copies for the input-gradient products (refreshed once per step, engine.refresh_transposes)
        self.data_t = torch.zeros(max(nt, 4), dtype=torch.float32, device=device)
        self.data = torch.zeros(n, dtype=torch.float32, device=device)
        # lo parts (x - tf32(x)) of both weight buffers for the tensor-core products (registered with the library in
        # Engine.__init__, refreshed together with the transposes)
        self.data_lo = torch.zeros_like(self.data)
        self.data_t_lo = torch.zeros_like(self.data_t)
        self.grad = torch.zeros(n, dtype=torch.float32, device=device)
        self.grid = {}        # name -> tensor
        self.grid_grad = {}
        self.pl_gain = float(conf.train.sigmoid_gain_lv_start)

    # split layers of geo / bg0 share one reference layer
    def _ref_layers(self, name):
        """[(ref_index, [Layer, ...])]"""
        out, ls = [], self.nets[name]
        if name in ("geo", "bg0"):
            for i, L in enumerate(ls[:-2]):
                out.append((i, [L]))
            out.append((len(ls) - 2, ls[-2:]))
        else:
            out = [(i, [L]) for i, L in enumerate(ls)]
        return out

    def load_reference(self, P):
        host = np.zeros(self.n_mlp, dtype=np.float32)
        for name in NET_ORDER:
            for i, parts in self._ref_layers(name):
                W, b = np.asarray(P[name][i][0], np.float32), np.asarray(P[name][i][1], np.float32)
                for L in parts:
                    Wi = np.zeros((L.K, L.ldw), np.float32)
                    Wi[L.rowmap, :L.N] = W[:, L.col0:L.col0 + L.N]
                    host[L.w_off:L.w_off + L.K * L.ldw] = Wi.ravel()
                    host[L.b_off:L.b_off + L.N] = b[L.col0:L.col0 + L.N]
        host[self.gain_off] = np.asarray(P["geo_gain"], np.float32).ravel()[0]
        self.data.copy_(torch.from_numpy(host))
        self.pl_gain = float(np.asarray(P["pl_gain"]).ravel()[0])
        for k, v in P.get("grid", {}).items():
            if v is not None:
                self.grid[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(self.device)
                self.grid_grad[k] = torch.zeros_like(self.grid[k])

    def init_grid_on_device(self, shapes, std=1e-3, seed=313):
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed)
        for k, shp in shapes.items():
            if k not in self.grid:
                self.grid[k] = torch.randn(shp, dtype=torch.float32, device=self.device, generator=gen) * std
                self.grid_grad[k] = torch.zeros_like(self.grid[k])

    def export_reference(self, which="grad"):
        """{'geo.W0': ..., 'geo.b0': ..., 'geo_gain': ..., 'grid.voxel': ...} numpy, reference layout."""
        buf = (self.grad if which == "grad" else self.data).detach().cpu().numpy()
        out = {}
        for name in NET_ORDER:
            for i, parts in self._ref_layers(name):
                Nref = sum(L.N for L in parts)
                Kref = len(parts[0].rowmap)
                W = np.zeros((Kref, Nref), np.float32)
                b = np.zeros(Nref, np.float32)
                for L in parts:
                    Wi = buf[L.w_off:L.w_off + L.K * L.ldw].reshape(L.K, L.ldw)
                    W[:, L.col0:L.col0 + L.N] = Wi[L.rowmap, :L.N]
                    b[L.col0:L.col0 + L.N] = buf[L.b_off:L.b_off + L.N]
                out[f"{name}.W{i}"], out[f"{name}.b{i}"] = W, b
        out["geo_gain"] = buf[self.gain_off:self.gain_off + 1].copy()
        src = self.grid_grad if which == "grad" else self.grid
        for k, v in src.items():
            out[f"grid.{k}"] = v.detach().cpu().numpy()
        return out

    def zero_grad(self):
        self.grad.zero_()
        for v in self.grid_grad.values():
            v.zero_()

    def lo_of(self, ptr):
        """Address of the pre-split lo part (x - tf32(x)) of a weight operand inside `data` / `data_t`, else 0."""
        if not isinstance(ptr, int):
            return 0
        for buf, lo in ((self.data, self.data_lo), (self.data_t, self.data_t_lo)):
            off = ptr - buf.data_ptr()
            if 0 <= off < 4 * buf.numel():
                return lo.data_ptr() + off
        return 0

    def W(self, L, row=0):
        return self.data.data_ptr() + 4 * (L.w_off + row * L.ldw)

    def b(self, L):
        return self.data.data_ptr() + 4 * L.b_off

    def Bt(self, L, row0=0):
        """Operand B(k, n) = W[row0 + n, k] of the input-gradient products as (pointer, b_rs, b_cs): the transposed copy
        (contiguous along n: TMA-friendly 128-byte rows) when the view is 16-byte aligned, else W itself."""
        if row0 % 4 == 0:
            return self.data_t.data_ptr() + 4 * (L.t_off + row0), L.ldt, 1
        sk = self.skip_t.get(id(L))
        if sk is not None and sk[0] == row0:
            return self.data_t.data_ptr() + 4 * sk[2], sk[3], 1
        return self.W(L, row0), 1, L.ldw

    def gW(self, L, row=0):
        return self.grad.data_ptr() + 4 * (L.w_off + row * L.ldw)

    def gb(self, L):
        return self.grad.data_ptr() + 4 * L.b_off


_ENGINES = {}


def get_engine(conf, **kw):
    """One Engine (parameters + scratch buffers) per configuration object, created on first use."""
    e = _ENGINES.get(id(conf))
    if e is None:
        e = _ENGINES[id(conf)] = Engine(conf, **kw)
    return e


def P_(t, off=0):
    """device address of element `off` of tensor t"""
    return t.data_ptr() + 4 * off


class Engine:
    def __init__(self, conf, device="cuda", world_size=1, process_group=None, grid_exchange="auto"):
        if not torch.cuda.is_available():
            raise _lib.NdjirError("ndjir_b200 needs a CUDA device (there is no CPU fallback)")
        _lib.lib()   # raises if the CUDA library is missing
        self.conf, self.device = conf, torch.device(device)
        self.params = ParamStore(conf, self.device)
        self.world_size, self.pg = world_size, process_group
        # multi-GPU exchange of the grid gradient: "dense" all-reduce of the table gradient, or "sparse" all-gather of
        # the per-sample scatter inputs followed by a replicated scatter (parallel.allgather_rows); auto = sparse for
        # tables above 256 MB (the 2 GiB voxel grid), dense for small ones
        self.grid_exchange = grid_exchange
        self._gather_cache = {}
        g = conf.geometric_network
        self.Df = g.feature_size
        self.Dg = grid_channels(conf)
        self.din = pe_dim(3, g.pe_bands) + self.Dg
        self.npe = pe_dim(3, g.pe_bands)
        self.ld0 = r4(self.din)
        self.LDO = r4(self.Df + 6)
        self.skip = g.skip_layers[0] if len(g.skip_layers) else -1
        assert len(g.skip_layers) <= 1 and g.geometric_init and not g.voxel.use_ste
        assert conf.diffuse_brdf.entangle and conf.specular_brdf.sampling == "importance"
        assert conf.background_modeling
        self.cskip = 1.0 / math.sqrt(2.0) if g.use_inv_square else 1.0
        self._bufs = {}
        self._graphs = {}
        self._reserve = 0
        self.one = torch.ones(4, dtype=torch.float32, device=self.device)
        self.rad = float(conf.renderer.bounding_sphere_radius)
        self.debug = {}
        self.profile, self.prof_events, self.n_launches = False, [], 0

    # ------------------------------------------------------------------------------------------------
    def stream(self):
        return torch.cuda.current_stream().cuda_stream

    def call(self, name, *args):
        self.n_launches += 1
        _lib.call(name, *args, self.stream())

    def buf(self, name, rows, cols, zero=False, dtype=torch.float32):
        key = (name, cols, dtype)
        t = self._bufs.get(key)
        if t is None or t.shape[0] < rows:
            t = torch.empty((max(rows, self._reserve), cols), dtype=dtype, device=self.device)
            self._bufs[key] = t
        if zero:
            t[:rows].zero_()
        return t

    def gemm(self, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, epi, bias=0, alpha=1.0, out_scale=1.0, H=0, ldh=0,
             hscale=1.0, U=0, ldu=0, C2=0, ldc2=0, split_k=1):
        if self.profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.call("ndjir_gemm_presplit", M, N, K, A, a_rs, a_cs, B, self.params.lo_of(B), b_rs, b_cs, C, ldc, bias,
                  alpha, out_scale, 100.0, H, ldh, hscale, U, ldu, C2, ldc2, split_k, epi)
        if self.profile:
            e1.record()
            self.prof_events.append((e0, e1, 2.0 * M * N * K, f"epi{epi} {M}x{N}x{K}"))

    def wgrad(self, rows, K, N, A, lda, dZ, ldz, gW, ldw, gb=0):
        """gW (K,N) += A(rows,K)^T dZ(rows,N): split-K over the rows with atomic accumulation; gb (N) += column sums
        of dZ (the bias gradient) in the same call when given."""
        tiles = ((K + 127) // 128) * ((N + 127) // 128 if N > 32 else 1)
        split = max(1, min(rows // 512, (148 * 4) // tiles))
        if not gb:
            self.gemm(K, N, rows, A, 1, lda, dZ, ldz, 1, gW, ldw, EPI_ATOMIC, split_k=split)
            return
        if self.profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.call("ndjir_wgrad_bias", rows, K, N, A, lda, dZ, ldz, gW, ldw, gb, split)
        if self.profile:
            e1.record()
            self.prof_events.append((e0, e1, 2.0 * K * N * rows, f"epi5 {K}x{N}x{rows}"))

    def copy2d(self, rows, cols, dst, ld_dst, src, ld_src, rep=1, alpha=1.0, accum=0):
        self.call("ndjir_copy2d", rows, cols, dst, ld_dst, src, ld_src, rep, alpha, accum)

    def _sparse_grid(self):
        if self.world_size <= 1:
            return False
        if self.grid_exchange == "auto":
            big = max([v.numel() * 4 for v in self.params.grid.values()] + [0])
            return big > (256 << 20)
        return self.grid_exchange == "sparse"

    def _gathered(self, t, rows, name, cache=False):
        """(world*rows, c) concatenation over ranks of t[:rows].  Query points are cached per step (the same x feeds
        several scatters); gradient rows are not (their scratch buffers are reused)."""
        key = (t.data_ptr(), rows, t.shape[1])
        if cache and key in self._gather_cache:
            return self._gather_cache[key]
        out = self.buf(f"gather_{name}", rows * self.world_size, t.shape[1])[:rows * self.world_size]
        g = allgather_rows(out, t[:rows].contiguous(), self.pg)
        if cache:
            self._gather_cache[key] = g
        return g

    def _grid_scatter(self, kind, part, rows, x, arrays, with_feature=False):
        """Scatter into the grid gradient (grad_feature / gqgf / tv_bwd).  arrays: dense (rows, c) per-sample tensors in
        the C-ABI order between grad_feature and query."""
        if self._sparse_grid():
            arrays = [self._gathered(a, rows, f"{kind}_{i}") for i, a in enumerate(arrays)]
            x = self._gathered(x, rows, f"x_{len(self._gather_cache)}_{x.data_ptr() % 9973}", cache=True)
            rows = rows * self.world_size
        tail = [P_(self.params.grid[part])] if with_feature else []
        self._grid_call(kind, part, rows, P_(self.params.grid_grad[part]), *[P_(a) for a in arrays], P_(x), *tail)

    def refresh_transposes(self):
        """W^T copies used by every input-gradient product; call after the parameters change (once per step)."""
        ps = self.params
        for name in NET_ORDER:
            for L in ps.nets[name]:
                self.call("ndjir_transpose", L.K, L.N, ps.data_t.data_ptr() + 4 * L.t_off, L.ldt, ps.W(L), L.ldw)
                sk = ps.skip_t.get(id(L))
                if sk is not None:      # rows [row0, row0 + rows) of W -> their own (N, r4(rows)) transposed block
                    self.call("ndjir_transpose", sk[1], L.N, ps.data_t.data_ptr() + 4 * sk[2], sk[3], ps.W(L, sk[0]), L.ldw)
        # pre-split lo parts of W and W^T: the weight operand of every tensor-core product arrives as two TMA tiles
        self.call("ndjir_split_lo", ps.data.numel(), ps.data_lo, ps.data)
        self.call("ndjir_split_lo", ps.data_t.numel(), ps.data_t_lo, ps.data_t)

    # ------------------------------------------------------------------------------------------------
    # grid feature dispatch (python/network.py:120-151 query_on_grid)
    # ------------------------------------------------------------------------------------------------
    def _grid_parts(self):
        v = self.conf.geometric_network.voxel
        if v.type == "voxel":
            return [("voxel", v.feature_size, 0)]
        if v.type == "triplaneline":
            return [("triplane", 3 * v.feature_size, 0), ("triline", 3 * v.feature_size, 3 * v.feature_size)]
        return []

    def _grid_call(self, kind, part, rows, *args):
        """kind in query/grad_query/grad_feature/ggo/gqgf/tv/tv_bwd; args are pointers in the C-ABI order between
        n_points and the grid spec."""
        v = self.conf.geometric_network.voxel
        G, D = v.grid_size, v.feature_size
        mn, mx = [-1.0] * 3, [1.0] * 3     # PF defaults min=-1, max=1 (voxel_feature.py:147-148)
        spec = ([G, G, G], D) if part == "voxel" else (G, D)
        names = {
            "voxel": dict(query="ndjir_voxel_query_on_voxel", grad_query="ndjir_voxel_grad_query",
                          grad_feature="ndjir_voxel_grad_feature", ggo="ndjir_voxel_grad_query_grad_grad_output",
                          gqgf="ndjir_voxel_grad_query_grad_feature", tv="ndjir_tv_loss_on_voxel",
                          tv_bwd="ndjir_tv_loss_on_voxel_backward"),
            "triplane": dict(query="ndjir_triplane_query_on_triplane", grad_query="ndjir_triplane_grad_query",
                             grad_feature="ndjir_triplane_grad_feature",
                             ggo="ndjir_triplane_grad_query_grad_grad_output",
                             gqgf="ndjir_triplane_grad_query_grad_feature", tv="ndjir_tv_loss_on_triplane",
                             tv_bwd="ndjir_tv_loss_on_triplane_backward"),
            "triline": dict(query="ndjir_triline_query_on_triline", grad_query="ndjir_triline_grad_query",
                            grad_feature="ndjir_triline_grad_feature",
                            ggo="ndjir_triline_grad_query_grad_grad_output",
                            gqgf="ndjir_triline_grad_query_grad_feature", tv="ndjir_tv_loss_on_triline",
                            tv_bwd="ndjir_tv_loss_on_triline_backward"),
        }[part]
        tail = {"query": (0,), "grad_query": (1,), "grad_feature": (1,), "ggo": (0,), "gqgf": (), "tv": (),
                "tv_bwd": (int(self.conf.train.tv_sym_backward),)}[kind]
        self.call(names[kind], rows, *args, *spec, mn, mx, *tail)

    # ------------------------------------------------------------------------------------------------
    # geometric network (python/network.py:154-232)
    # ------------------------------------------------------------------------------------------------
    def geo_input(self, x, rows, A0, tag):
        """A0 (rows, ld0) = [PE(x) | grid features]"""
        g = self.conf.geometric_network
        self.call("ndjir_positional_encoding", rows, 3, g.pe_bands, P_(x), 3, 1, P_(A0), self.ld0)
        for part, width, off in self._grid_parts():
            tmp = self.buf(f"gq_{part}", rows, width)
            self._grid_call("query", part, rows, P_(tmp), P_(x), P_(self.params.grid[part]))
            self.copy2d(rows, width, P_(A0, self.npe + off), self.ld0, P_(tmp), width)
        if self.ld0 > self.din:
            self.call("ndjir_copy2d", rows, self.ld0 - self.din, P_(A0, self.din), self.ld0, P_(self.one), 0, 1, 0.0, 0)

    def geo_forward(self, x, rows, tag, store, want_feat=True, sdf_out=None, O=None):
        """Forward pass.  store=True keeps every layer input in buffers `{tag}_A{l}` for the backward passes.
        Returns (A list, sdf tensor)."""
        ps, net = self.params, self.params.nets["geo"]
        nl = len(net) - 2          # hidden layers (reference layers 0..L-2)
        A = [self.buf(f"{tag}_A0", rows, self.ld0)]
        self.geo_input(x, rows, A[0], tag)
        for l in range(nl):
            L = net[l]
            name = f"{tag}_A{l + 1}" if store else f"geo_pp{l % 2}"
            ldn = r4(net[l + 1].K)
            nxt = self.buf(name, rows, ldn)
            osc = self.cskip if (l + 1) == self.skip else 1.0
            self.gemm(rows, L.N, L.K, P_(A[l]), A[l].shape[1], 1, ps.W(L), L.ldw, 1, P_(nxt), ldn, EPI_SOFTPLUS,
                      bias=ps.b(L), out_scale=osc)
            if (l + 1) == self.skip:
                self.copy2d(rows, self.din, P_(nxt, L.N), ldn, P_(A[0]), self.ld0, alpha=self.cskip)
            A.append(nxt)
        Ls, Lf = net[-2], net[-1]
        sdf = sdf_out if sdf_out is not None else self.buf(f"{tag}_sdf", rows, 1)
        last = A[-1]
        self.gemm(rows, 1, Ls.K, P_(last), last.shape[1], 1, ps.W(Ls), Ls.ldw, 1, P_(sdf), 1, EPI_BIAS, bias=ps.b(Ls))
        if want_feat:
            self.gemm(rows, Lf.N, Lf.K, P_(last), last.shape[1], 1, ps.W(Lf), Lf.ldw, 1, P_(O), self.LDO, EPI_BIAS,
                      bias=ps.b(Lf))
        return A, sdf

    def geo_normal(self, x, rows, tag, A, nrm):
        """n = d sdf / d x by a reverse sweep (the nn.grad graph of renderer.py:52).  Keeps GZ[l] = dsdf/dz_l."""
        ps, net = self.params, self.params.nets["geo"]
        nl = len(net) - 2
        Ls = net[-2]
        GZ = [self.buf(f"{tag}_GZ{l}", rows, r4(net[l].N)) for l in range(nl)]
        Gin = self.buf(f"{tag}_Gin", rows, self.ld0, zero=True)
        c = self.cskip
        # top: gA_{L-1}[p,:] = w_sdf ; GZ[nl-1] = gA * s
        top = A[nl]
        self.gemm(rows, Ls.K, 1, P_(self.one), 0, 1, *ps.Bt(Ls), P_(GZ[nl - 1]), GZ[nl - 1].shape[1],
                  EPI_MUL_S, H=P_(top), ldh=top.shape[1])
        skip_wrote = False
        for l in range(nl - 1, 0, -1):
            L = net[l]
            is_skip = (l == self.skip)
            n_prev = net[l - 1].N
            self.gemm(rows, n_prev, L.N, P_(GZ[l]), GZ[l].shape[1], 1, *ps.Bt(L), P_(GZ[l - 1]),
                      GZ[l - 1].shape[1], EPI_MUL_S, alpha=(c if is_skip else 1.0), H=P_(A[l]), ldh=A[l].shape[1],
                      hscale=(1.0 / c if is_skip else 1.0))
            if is_skip:
                self.gemm(rows, self.din, L.N, P_(GZ[l]), GZ[l].shape[1], 1, *ps.Bt(L, n_prev), P_(Gin),
                          self.ld0, EPI_BIAS, alpha=c)
                skip_wrote = True
        L0 = net[0]
        self.gemm(rows, self.din, L0.N, P_(GZ[0]), GZ[0].shape[1], 1, *ps.Bt(L0), P_(Gin), self.ld0,
                  EPI_ACCUM if skip_wrote else EPI_BIAS)
        g = self.conf.geometric_network
        self.call("ndjir_positional_encoding_grad_input", rows, 3, g.pe_bands, P_(A[0]), self.ld0, P_(Gin), self.ld0,
                  P_(nrm), 3, 0)
        for part, width, off in self._grid_parts():
            tmp = self.buf(f"gg_{part}", rows, width)
            self.copy2d(rows, width, P_(tmp), width, P_(Gin, self.npe + off), self.ld0)
            self._grid_call("grad_query", part, rows, P_(nrm), P_(tmp), P_(x), P_(self.params.grid[part]))
        return GZ, Gin

    def geo_normal_adjoint(self, x, rows, tag, A, GZ, Gin, nbar):
        """Reverse of geo_normal given nbar = dL/dn (P,3): accumulates the weight gradients of the normal pass,
        the grid gradient through d/dq, and returns Z2[l] = second-order contribution to dL/dz_l."""
        ps, net = self.params, self.params.nets["geo"]
        nl = len(net) - 2
        g = self.conf.geometric_network
        c = self.cskip
        Ghat = [self.buf(f"{tag}_Gh0", rows, self.ld0, zero=True)]
        self.call("ndjir_positional_encoding_grad_input_adjoint", rows, 3, g.pe_bands, P_(A[0]), self.ld0, P_(nbar), 3,
                  P_(Ghat[0]), self.ld0)
        for part, width, off in self._grid_parts():
            tmp = self.buf(f"gg_{part}", rows, width)       # g_in grid columns (dense)
            self.copy2d(rows, width, P_(tmp), width, P_(Gin, self.npe + off), self.ld0)
            tmp2 = self.buf(f"ggo_{part}", rows, width)
            self._grid_call("ggo", part, rows, P_(tmp2), P_(nbar), P_(x), P_(self.params.grid[part]))
            self.copy2d(rows, width, P_(Ghat[0], self.npe + off), self.ld0, P_(tmp2), width)
            self._grid_scatter("gqgf", part, rows, x, [nbar, tmp])
        Z2 = []
        for l in range(nl):
            L = net[l]
            ldn = r4(net[l + 1].K)
            nxt = self.buf(f"{tag}_Gh{l + 1}", rows, ldn)
            z2 = self.buf(f"{tag}_Z2{l}", rows, r4(L.N))
            is_skip = (l + 1) == self.skip
            Gh = Ghat[l]
            self.gemm(rows, L.N, L.K, P_(Gh), Gh.shape[1], 1, ps.W(L), L.ldw, 1, P_(z2), z2.shape[1], EPI_ADJ,
                      out_scale=(c if is_skip else 1.0), H=P_(A[l + 1]), ldh=A[l + 1].shape[1],
                      hscale=(1.0 / c if is_skip else 1.0), U=P_(GZ[l]), ldu=GZ[l].shape[1], C2=P_(nxt), ldc2=ldn)
            if is_skip:
                self.copy2d(rows, self.din, P_(nxt, L.N), ldn, P_(Ghat[0]), self.ld0, alpha=c)
            self.wgrad(rows, L.K, L.N, P_(Gh), Gh.shape[1], P_(GZ[l]), GZ[l].shape[1], ps.gW(L), L.ldw)
            Ghat.append(nxt)
            Z2.append(z2)
        Ls = net[-2]
        top = Ghat[nl]
        # g w_sdf[k] += sum_p Ghat_top[p,k]
        self.gemm(Ls.K, 1, rows, P_(top), 1, top.shape[1], P_(self.one), 0, 1, ps.gW(Ls), Ls.ldw, EPI_ATOMIC,
                  split_k=max(1, min(rows // 512, 148)))
        return Z2

    def geo_backward(self, x, rows, tag, A, dsdf, dO, Z2):
        """Standard reverse sweep.  dsdf (rows,1) or None, dO (rows, LDO) holds dL/dfeature in columns 0:Df,
        Z2 = second-order terms from geo_normal_adjoint or None.  Scatters the grid gradient."""
        ps, net = self.params, self.params.nets["geo"]
        nl = len(net) - 2
        Ls, Lf = net[-2], net[-1]
        c = self.cskip
        last = A[nl]
        pp = [self.buf("geo_dz0", rows, r4(self.Df)), self.buf("geo_dz1", rows, r4(self.Df))]
        # last layer
        self.wgrad(rows, Lf.K, Lf.N, P_(last), last.shape[1], P_(dO), self.LDO, ps.gW(Lf), Lf.ldw, gb=ps.gb(Lf))
        cur = pp[0]
        ldc = cur.shape[1]
        self.gemm(rows, Lf.K, Lf.N, P_(dO), self.LDO, 1, *ps.Bt(Lf), P_(cur), ldc, EPI_MUL_S, H=P_(last),
                  ldh=last.shape[1], U=(P_(Z2[nl - 1]) if Z2 else 0), ldu=(Z2[nl - 1].shape[1] if Z2 else 0))
        if dsdf is not None:
            self.wgrad(rows, Ls.K, 1, P_(last), last.shape[1], P_(dsdf), 1, ps.gW(Ls), Ls.ldw)
            self.call("ndjir_colsum", rows, 1, ps.gb(Ls), P_(dsdf), 1, 1.0)
            self.gemm(rows, Ls.K, 1, P_(dsdf), 1, 1, *ps.Bt(Ls), P_(cur), ldc, EPI_MUL_S, H=P_(last),
                      ldh=last.shape[1], U=P_(cur), ldu=ldc)
        dgrid = self.buf("geo_dgrid", rows, max(self.Dg, 1)) if self.Dg else None
        skip_wrote = False
        for l in range(nl - 1, -1, -1):
            L = net[l]
            Al = A[l]
            self.wgrad(rows, L.K, L.N, P_(Al), Al.shape[1], P_(cur), ldc, ps.gW(L), L.ldw, gb=ps.gb(L))
            if l > 0:
                is_skip = (l == self.skip)
                n_prev = net[l - 1].N
                nxt = pp[1] if cur is pp[0] else pp[0]
                self.gemm(rows, n_prev, L.N, P_(cur), ldc, 1, *ps.Bt(L), P_(nxt), nxt.shape[1], EPI_MUL_S,
                          alpha=(c if is_skip else 1.0), H=P_(Al), ldh=Al.shape[1],
                          hscale=(1.0 / c if is_skip else 1.0), U=(P_(Z2[l - 1]) if Z2 else 0),
                          ldu=(Z2[l - 1].shape[1] if Z2 else 0))
                if is_skip and self.Dg:
                    self.gemm(rows, self.Dg, L.N, P_(cur), ldc, 1, *ps.Bt(L, n_prev + self.npe), P_(dgrid),
                              self.Dg, EPI_BIAS, alpha=c)
                    skip_wrote = True
                cur, ldc = nxt, nxt.shape[1]
            elif self.Dg:
                self.gemm(rows, self.Dg, L.N, P_(cur), ldc, 1, *ps.Bt(L, self.npe), P_(dgrid), self.Dg,
                          EPI_ACCUM if skip_wrote else EPI_BIAS)
        for part, width, off in self._grid_parts():
            tmp = self.buf(f"gg_{part}", rows, width)
            self.copy2d(rows, width, P_(tmp), width, P_(dgrid, off), self.Dg)
            self._grid_scatter("grad_feature", part, rows, x, [tmp])

    # ------------------------------------------------------------------------------------------------
    # generic softplus MLP (heads): forward keeps layer inputs, backward accumulates weight gradients
    # ------------------------------------------------------------------------------------------------
    def mlp_forward(self, name, tag, X, ldx, rows, outs):
        """outs: list of (ptr, ldc) for the (possibly split) last reference layer."""
        ps, net = self.params, self.params.nets[name]
        n_last = len(outs)
        nh = len(net) - n_last
        acts = []
        Aptr, lda = X, ldx
        for l in range(nh):
            L = net[l]
            nxt = self.buf(f"{tag}_h{l}", rows, r4(L.N))
            self.gemm(rows, L.N, L.K, Aptr, lda, 1, ps.W(L), L.ldw, 1, P_(nxt), nxt.shape[1], EPI_SOFTPLUS,
                      bias=ps.b(L))
            acts.append(nxt)
            Aptr, lda = P_(nxt), nxt.shape[1]
        for (optr, ldo), L in zip(outs, net[nh:]):
            self.gemm(rows, L.N, L.K, Aptr, lda, 1, ps.W(L), L.ldw, 1, optr, ldo, EPI_BIAS, bias=ps.b(L))
        return acts

    def mlp_backward(self, name, tag, X, ldx, rows, acts, douts, dX=0, lddx=0, accum_dx=False, dx_cols=None):
        """douts: list of (ptr, ld) matching the last (split) layers.  dX (rows, dx_cols) (+)= input gradient."""
        ps, net = self.params, self.params.nets[name]
        n_last = len(douts)
        nh = len(net) - n_last
        wmax = max(r4(L.N) for L in net[:nh]) if nh else 4
        pp = [self.buf("mlp_dz0", rows, wmax), self.buf("mlp_dz1", rows, wmax)]   # keyed by (name, width)
        lastA, lda = (P_(acts[-1]), acts[-1].shape[1]) if nh else (X, ldx)
        cur = pp[0]
        first = True
        for (dptr, ldd), L in zip(douts, net[nh:]):
            self.wgrad(rows, L.K, L.N, lastA, lda, dptr, ldd, ps.gW(L), L.ldw, gb=ps.gb(L))
            if nh:
                self.gemm(rows, L.K, L.N, dptr, ldd, 1, *ps.Bt(L), P_(cur), wmax, EPI_MUL_S, H=lastA, ldh=lda,
                          U=(0 if first else P_(cur)), ldu=(0 if first else wmax))
            elif dX:
                self.gemm(rows, dx_cols or L.K, L.N, dptr, ldd, 1, *ps.Bt(L), dX, lddx,
                          EPI_ACCUM if (accum_dx or not first) else EPI_BIAS)
            first = False
        for l in range(nh - 1, -1, -1):
            L = net[l]
            Aptr, lda = (P_(acts[l - 1]), acts[l - 1].shape[1]) if l > 0 else (X, ldx)
            self.wgrad(rows, L.K, L.N, Aptr, lda, P_(cur), wmax, ps.gW(L), L.ldw, gb=ps.gb(L))
            if l > 0:
                nxt = pp[1] if cur is pp[0] else pp[0]
                self.gemm(rows, L.K, L.N, P_(cur), wmax, 1, *ps.Bt(L), P_(nxt), wmax, EPI_MUL_S, H=Aptr,
                          ldh=lda)
                cur = nxt
            elif dX:
                self.gemm(rows, dx_cols or L.K, L.N, P_(cur), wmax, 1, *ps.Bt(L), dX, lddx,
                          EPI_ACCUM if accum_dx else EPI_BIAS)

    # ------------------------------------------------------------------------------------------------
    # sample_points (python/sampler.py:256-299)
    # ------------------------------------------------------------------------------------------------
    def sample_points(self, camloc, raydir, stratified_sample, background_sample, mask_sum=None, debug=False):
        """camloc (B,3), raydir (B,R,3), stratified_sample (B,R,N0,1), background_sample (B,R,Nb+1,1) device fp32.
        Returns x_fg (B,R,N,3), t_fg (B,R,N+1,1), x_bg (B,R,Nb,4), t_bg (B,R,Nb+1,1), mask (B,R,1,1)."""
        r = self.conf.renderer
        B, R, _ = raydir.shape
        NR = B * R
        if not getattr(self, "_weights_synced", False):
            self.refresh_transposes()   # standalone call: the registered lo copies of the weights must be current
        N0, M, U, Nb = r.n_samples0, r.n_samples1, r.n_upsamples, r.n_bg_samples
        N = N0 + U * M
        tn, tf, nh = (self.buf(k, NR, 1) for k in ("t_near", "t_far", "n_hits"))
        mask = self.buf("mask", NR, 1)
        if r.t_near_far_method == "intersect_with_aabb":
            self.call("ndjir_ray_aabb_intersection", NR, P_(tn), P_(tf), P_(nh), P_(camloc), P_(raydir), B, R,
                      [-self.rad] * 3, [self.rad] * 3)
        elif r.t_near_far_method == "intersect_with_r_sphere":
            self.call("ndjir_ray_sphere_intersection", NR, P_(tn), P_(tf), P_(nh), P_(camloc), P_(raydir), B, R,
                      self.rad)
        else:
            raise NotImplementedError(r.t_near_far_method)
        self.call("ndjir_hit_mask", NR, P_(nh), P_(mask), P_(mask_sum) if mask_sum is not None else 0)
        ld = N + 1
        cur = self.buf("t_a", NR, N + 1)             # sorted distances, merged in place round by round
        sdf_m = self.buf("smp_sdf_m", NR, N)           # their SDF values, carried along (never re-evaluated)
        pend_t = self.buf("smp_t_pend", NR, max(N0, M))
        self.call("ndjir_stratified_dists", NR, N0, P_(pend_t), P_(tn), P_(tf), P_(stratified_sample))
        Nt, Mp = 0, N0
        dbg = []
        self._reserve = NR * max(N0, M)      # size the sampler's scratch for the largest evaluation up front
        x = self.buf("smp_x", NR * max(N0, M), 3)
        sdf_p = self.buf("smp_sdf", NR * max(N0, M), 1)
        for u in range(U + 1):
            # SDF of the pending samples only: the stratified ones, then the M new ones of each round.  The reference
            # re-evaluates every current sample per round (sampler.py:190-192), 352 evaluations per ray instead of
            # 112, with identical values.
            last = u == U
            if not last:
                self.call("ndjir_ray_points", NR, Mp, R, P_(x), P_(camloc), P_(raydir), P_(pend_t), Mp)
                self.geo_forward(x, NR * Mp, "smp", store=False, want_feat=False, sdf_out=sdf_p)
            gain = float(r.sampling_sigmoid_gain * 2 ** u)
            tnew = self.buf(f"smp_tnew{u % 2}", NR, M)
            idx = self.buf(f"smp_idx{u}", NR, M, dtype=torch.int32) if (debug and not last) else None
            # the last call only merges the final M samples (their SDF is not needed: sdf_p is stale and unused)
            self.call("ndjir_importance_round_incremental", NR, Nt, Mp, 0 if last else M, P_(cur), ld, P_(sdf_m), N,
                      P_(pend_t), P_(sdf_p), P_(tn), P_(tf), gain, P_(tnew), idx.data_ptr() if idx is not None else 0)
            Nt += Mp
            if debug:
                if dbg:
                    dbg[-1]["t_out"] = cur[:NR, :Nt].clone()
                if not last:
                    dbg.append(dict(t_in=cur[:NR, :Nt].clone(), sdf=sdf_m[:NR, :Nt].clone(), t_new=tnew[:NR].clone(),
                                    idx=idx[:NR].clone()))
            pend_t, Mp = tnew, M
        self._reserve = 0
        # t_fg = concat(t, t_far)
        self.copy2d(NR, 1, P_(cur, N), ld, P_(tf), 1)
        t_fg = cur[:NR].reshape(B, R, N + 1, 1)
        x_fg = self.buf("x_fg", NR * N, 3)[:NR * N].view(B, R, N, 3)
        self.call("ndjir_ray_points", NR, N, R, P_(x_fg), P_(camloc), P_(raydir), P_(cur), ld)
        t_bg = self.buf("t_bg", NR, Nb + 1)[:NR].view(B, R, Nb + 1, 1)
        x_bg = self.buf("x_bg", NR * Nb, 4)[:NR * Nb].view(B, R, Nb, 4)
        self.call("ndjir_background_samples", NR, Nb, R, P_(camloc), P_(raydir), P_(tf), P_(mask), P_(background_sample),
                  self.rad, P_(t_bg), P_(x_bg))
        self.debug["sampler"] = dbg
        return x_fg, t_fg, x_bg, t_bg, mask[:NR].reshape(B, R, 1, 1)

    # ------------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole step
    # ------------------------------------------------------------------------------------------------
    def train_step_graphed(self, camloc, raydir, color_gt, rnd, cos_anneal_ratio=0.0):
        """train_step(zero_grad=True) captured ONCE in a CUDA graph and replayed: the step is ~2200 kernel launches
        (31 ms of host time to enqueue for 45 ms of device time, and the gaps between dependent small kernels cost
        ~1 ms per step - tools/exp/graph_step.py).  Inputs (device or pinned-host tensors) are copied into static
        device buffers, the returned loss tensor is the graph's static output.  The host scalars baked into the kernel
        arguments (cos_anneal_ratio, the light-visibility gain) and the input shapes key the cache: a change captures
        a new graph.  Single-process only (the NCCL exchanges of the ray-sharded mode stay on the eager path)."""
        if self.world_size > 1:
            return self.train_step(camloc, raydir, color_gt, rnd, cos_anneal_ratio=cos_anneal_ratio)
        inputs = {"camloc": camloc, "raydir": raydir, "color_gt": color_gt, **rnd}
        key = (float(cos_anneal_ratio), float(self.params.pl_gain),
               tuple((k, tuple(v.shape)) for k, v in sorted(inputs.items())))
        st = self._graphs.get(key)
        if st is None:
            static = {k: torch.empty(v.shape, dtype=torch.float32, device=self.device) for k, v in inputs.items()}
            for k, v in inputs.items():
                static[k].copy_(v, non_blocking=True)

            def run():
                rest = {k: v for k, v in static.items() if k not in ("camloc", "raydir", "color_gt")}
                return self.train_step(static["camloc"], static["raydir"], static["color_gt"], rest,
                                       cos_anneal_ratio=cos_anneal_ratio)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):       # eager warm-up: sizes every scratch buffer, sets kernel attributes
                run()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = run()
            st = self._graphs[key] = (graph, static, out)
            if len(self._graphs) > 4:           # graphs pin their scratch memory: keep only the most recent ones
                self._graphs.pop(next(iter(self._graphs)))
        graph, static, out = st
        for k, v in inputs.items():
            static[k].copy_(v, non_blocking=True)
        graph.replay()
        return out

    # ------------------------------------------------------------------------------------------------
    # total_loss forward + backward (python/loss.py:27-192 over python/renderer.py:32-209)
    # ------------------------------------------------------------------------------------------------
    def train_step(self, camloc, raydir, color_gt, rnd, cos_anneal_ratio=0.0, samples=None, zero_grad=True,
                   backward=True, keep=False, inference=False):
        """One loss.forward() + loss.backward().  `rnd` holds the explicit random tensors (scene.make_randoms) on
        the device.  Returns the (N_LOSSES,) device tensor of loss terms; gradients accumulate in self.params.grad /
        grid_grad (all-reduced over the process group when world_size > 1)."""
        conf, ps = self.conf, self.params
        r, tr = conf.renderer, conf.train
        B, R, _ = raydir.shape
        NR = B * R
        N = r.n_samples0 + r.n_upsamples * r.n_samples1
        Nb, nt = r.n_bg_samples, r.n_thetas
        M = nt * 2 * nt
        P = NR * N
        Df, LDO = self.Df, self.LDO
        S = N + Nb
        if zero_grad:
            ps.zero_grad()
        self.refresh_transposes()       # the normal pass (forward half) already needs W^T
        self._weights_synced = True
        self._gather_cache = {}
        losses = self.buf("losses", 1, 16, zero=True)
        scal = self.buf("scalars", 1, 8, zero=True)      # [mask_sum, inv_denorm]
        mask_sum, inv_denorm = P_(scal, 0), P_(scal, 1)
        if samples is None:
            x_fg, t_fg, x_bg, t_bg, mask = self.sample_points(camloc, raydir, rnd["stratified"], rnd["background"],
                                                              mask_sum=scal)
        else:
            x_fg, t_fg, x_bg, t_bg, mask = samples
            scal[0, 0] = mask.sum()
        if self.world_size > 1:
            allreduce_mask_sum(scal[0, 0:1], self.pg)
        self.call("ndjir_loss_inv_denorm", mask_sum, N, inv_denorm)
        inv_rays = 1.0 / (NR * self.world_size)
        x_fg = x_fg.reshape(P, 3)
        maskv = mask.reshape(NR)
        # ---------------- geometric network + normal ----------------
        O = self.buf("O", P, LDO)
        A, sdf = self.geo_forward(x_fg, P, "main", store=True, O=O)
        nrm = self.buf("nrm", P, 3)
        GZ, Gin = self.geo_normal(x_fg, P, "main", A, nrm)
        self.copy2d(P, 3, P_(O, Df), LDO, P_(x_fg), 3)
        self.copy2d(P, 3, P_(O, Df + 3), LDO, P_(nrm), 3)
        # ---------------- NeuS alpha, background, compositing ----------------
        alpha_fg = self.buf("alpha_fg", P, 1)
        gain_p = ps.data.data_ptr() + 4 * ps.gain_off
        self.call("ndjir_neus_alpha_forward", P, N, P_(alpha_fg), P_(sdf), P_(nrm), 3, P_(raydir), P_(t_fg), gain_p,
                  float(cos_anneal_ratio))
        bgc = conf.background_network
        rows_bg = NR * Nb
        Xbg0 = self.buf("Xbg0", rows_bg, r4(pe_dim(4, bgc.pe_bands0)))
        self.call("ndjir_positional_encoding", rows_bg, 4, bgc.pe_bands0, P_(x_bg), 4, 1, P_(Xbg0), Xbg0.shape[1])
        Dfb = bgc.feature_size0
        nvpe = pe_dim(3, bgc.pe_bands1)
        ldb1 = r4(Dfb + 4 + 3 + nvpe)
        Xbg1 = self.buf("Xbg1", rows_bg, ldb1)
        dens = self.buf("bg_dens", rows_bg, 1)
        acts_bg0 = self.mlp_forward("bg0", "bg0", P_(Xbg0), Xbg0.shape[1], rows_bg, [(P_(dens), 1), (P_(Xbg1), ldb1)])
        self.copy2d(rows_bg, 4, P_(Xbg1, Dfb), ldb1, P_(x_bg), 4)
        view = self.buf("view", NR, 3)
        self.copy2d(NR, 3, P_(view), 3, P_(raydir), 3, alpha=-1.0)
        self.copy2d(rows_bg, 3, P_(Xbg1, Dfb + 4), ldb1, P_(view), 3, rep=Nb)
        vpe_bg = self.buf("vpe_bg", NR, r4(nvpe))
        self.call("ndjir_positional_encoding", NR, 3, bgc.pe_bands1, P_(view), 3, 1, P_(vpe_bg), vpe_bg.shape[1])
        self.copy2d(rows_bg, nvpe, P_(Xbg1, Dfb + 7), ldb1, P_(vpe_bg), vpe_bg.shape[1], rep=Nb)
        bgraw = self.buf("bg_raw", rows_bg, 4)
        acts_bg1 = self.mlp_forward("bg1", "bg1", P_(Xbg1), ldb1, rows_bg, [(P_(bgraw), 4)])
        alpha_bg = self.buf("alpha_bg", rows_bg, 1)
        self.call("ndjir_bg_alpha_forward", rows_bg, Nb, P_(alpha_bg), P_(dens), 1, P_(t_bg))
        w = self.buf("w", NR, S)
        T = self.buf("T", NR, S)
        self.call("ndjir_composite_forward", NR, N, Nb, P_(alpha_fg), P_(maskv), P_(alpha_bg), P_(w), P_(T))
        colbg = self.buf("colbg", NR, 3)
        self.call("ndjir_bg_color_forward", NR, Nb, P_(w, N), S, P_(bgraw), 4, P_(colbg))
        # ---------------- pixel quantities ----------------
        pix = self.buf("pix", NR, LDO)
        self.call("ndjir_volume_render_forward", NR, N, Df + 6, P_(w), S, P_(O), LDO, P_(pix), LDO)
        nhat = self.buf("nhat", NR, 3)
        self.call("ndjir_pixel_normal_forward", NR, P_(pix, Df + 3), LDO, float(r.eps_normal), P_(nhat))
        # ---------------- per-sample heads ----------------
        RAW = self.buf("RAW", P, 16, zero=inference)     # inference leaves column 13 (perturbed base colour) unset
        acts = {}
        acts["bc"] = self.mlp_forward("bc", "bc", P_(O), LDO, P, [(P_(RAW, 0), 16)])
        acts["ii"] = self.mlp_forward("ii", "ii", P_(O), LDO, P, [(P_(RAW, 3), 16)])
        acts["ro"] = self.mlp_forward("ro", "ro", P_(O), LDO, P, [(P_(RAW, 4), 16)])
        acts["sp"] = self.mlp_forward("sp", "sp", P_(O), LDO, P, [(P_(RAW, 6), 16)])
        plc = conf.photogrammetric_light_network
        npl = pe_dim(3, plc.pe_bands)
        ldpl = r4(Df + 6 + npl + 1)
        Xpl = self.buf("Xpl", P, ldpl)
        self.copy2d(P, Df + 6, P_(Xpl), ldpl, P_(O), LDO)
        vpe_pl = self.buf("vpe_pl", NR, r4(npl))
        self.call("ndjir_positional_encoding", NR, 3, plc.pe_bands, P_(view), 3, 1, P_(vpe_pl), vpe_pl.shape[1])
        self.copy2d(P, npl, P_(Xpl, Df + 6), ldpl, P_(vpe_pl), vpe_pl.shape[1], rep=N)
        self.call("ndjir_inv_sq_dist", P, R * N, P_(x_fg), P_(camloc), P_(Xpl, Df + 6 + npl), ldpl)
        acts["pl"] = self.mlp_forward("pl", "pl", P_(Xpl), ldpl, P, [(P_(RAW, 12), 16)])
        # ---------------- perturbed colour branch (renderer.py:187-193) ----------------
        # It only feeds the base-colour prior of the loss: image rendering (inference=True, forward only) skips the
        # second geometric-network evaluation and reads a zero base colour there.
        assert not (inference and backward), "inference=True is forward only"
        if not inference:
            G = conf.geometric_network.voxel.grid_size
            x_ptb = self.buf("x_ptb", P, 3)
            self.copy2d(P, 3, P_(x_ptb), 3, P_(x_fg), 3)
            self.copy2d(P, 3, P_(x_ptb), 3, P_(rnd["perturb"]), 3, alpha=math.sqrt(3) * 2 * self.rad / G, accum=1)
            Op = self.buf("O_ptb", P, LDO)
            Ap, _ = self.geo_forward(x_ptb, P, "ptb", store=True, O=Op)
            self.copy2d(P, 3, P_(Op, Df), LDO, P_(x_ptb), 3)
            acts["bcp"] = self.mlp_forward("bc", "bcp", P_(Op), LDO, P, [(P_(RAW, 13), 16)])
        # ---------------- material attributes + per-sample losses ----------------
        ro_c, sp_c = conf.roughness_network, conf.specular_reflectance_network
        cfg10 = [ro_c.lower_bound, ro_c.prior_value, sp_c.prior_value, sp_c.upper_bound_scale, ps.pl_gain,
                 tr.eikonal_weight, tr.base_color_prior_weight, tr.roughness_prior_weight,
                 tr.specular_reflectance_prior_weight, float(tr.base_color_prior_sym_backward)]
        ATT = self.buf("ATT", P, 12)
        self.call("ndjir_sample_attributes_forward", P, N, P_(RAW), P_(ATT), P_(nrm), 3, P_(maskv), cfg10, P_(losses))
        attpix = self.buf("attpix", NR, 12)
        self.call("ndjir_volume_render_forward", NR, N, 12, P_(w), S, P_(ATT), 12, P_(attpix), 12)
        # ---------------- light directions, environment light, soft visibility ----------------
        dirs_u = self.buf("dirs_u", NR * M, 3)
        dirs_s = self.buf("dirs_s", NR * M, 3)
        self.call("ndjir_sample_uniform_directions", NR * M, P_(dirs_u), P_(nhat), P_(rnd["diffuse_cdf_the"]),
                  P_(rnd["diffuse_cdf_phi"]), NR, M, nt, 2 * nt, 0.0)
        rho = self.buf("rho_pix", NR, 1)
        self.copy2d(NR, 1, P_(rho), 1, P_(attpix, 1), 12)
        self.call("ndjir_sample_importance_directions", NR * M, P_(dirs_s), P_(nhat), P_(rnd["specular_cdf_the"]),
                  P_(rnd["specular_cdf_phi"]), P_(rho), NR, M, nt, 2 * nt, 0.0)
        elc, svc = conf.environment_light_network, conf.soft_visibility_light_network
        rows_d = NR * 2 * M
        nel = pe_dim(3, elc.pe_bands)
        Xel = self.buf("Xel", rows_d, r4(nel))
        ldel = Xel.shape[1]
        # rows (set, r, j): the diffuse set first, then the specular set
        for s_, dirs in enumerate((dirs_u, dirs_s)):
            self.call("ndjir_positional_encoding", NR * M, 3, elc.pe_bands, P_(dirs), 3, 1, P_(Xel, s_ * NR * M * ldel),
                      ldel)
        elraw = self.buf("el_raw", rows_d, 4)
        acts["el"] = self.mlp_forward("el", "el", P_(Xel), ldel, rows_d, [(P_(elraw), 4)])
        nsv = pe_dim(3, svc.pe_bands)
        assert nsv == nel
        ldsv = r4(Df + 6 + nsv)
        Xsv = self.buf("Xsv", rows_d, ldsv)
        for s_ in range(2):
            o_ = s_ * NR * M * ldsv
            self.copy2d(NR * M, Df + 3, P_(Xsv, o_), ldsv, P_(pix), LDO, rep=M)            # f_pix | x_pix
            self.copy2d(NR * M, 3, P_(Xsv, o_ + Df + 3), ldsv, P_(nhat), 3, rep=M)         # nhat
        self.copy2d(rows_d, nsv, P_(Xsv, Df + 6), ldsv, P_(Xel), ldel)                     # PE(omega)
        svraw = self.buf("sv_raw", rows_d, 4)
        acts["sv"] = self.mlp_forward("sv", "sv", P_(Xsv), ldsv, rows_d, [(P_(svraw), 4)])
        # ---------------- shading + colour loss ----------------
        cfg5 = [r.eps_dot, conf.specular_brdf.weight, inv_rays, 1.0, 0.0 if tr.rgb_loss == "l1" else 1.0]
        color = self.buf("color", NR, 3)
        self.call("ndjir_shade_forward", NR, M, P_(nhat), P_(attpix), P_(raydir), P_(dirs_u), P_(dirs_s), P_(elraw), 4,
                  P_(svraw), 4, P_(colbg), P_(color_gt), cfg5, P_(color), P_(losses))
        # ---------------- TV loss ----------------
        tv_on = conf.geometric_network.voxel.type != "none" and tr.tv_weight > 0
        if tv_on:
            for part, width, off in self._grid_parts():
                tvb = self.buf(f"tv_{part}", P, width)
                self._grid_call("tv", part, P, P_(tvb), P_(x_fg), P_(ps.grid[part]))
                self.call("ndjir_masked_sum", P, N, width, P_(tvb), P_(maskv), P_(losses, 3))
        self.call("ndjir_finalize_losses", P_(losses), mask_sum, N, inv_rays, tr.eikonal_weight,
                  tr.tv_weight if tv_on else 0.0, tr.base_color_prior_weight, tr.roughness_prior_weight,
                  tr.specular_reflectance_prior_weight)
        if self.world_size > 1:
            allreduce_mask_sum(losses[0, :N_LOSSES], self.pg)     # report the loss of the union of all ranks' rays
        if keep:
            self.debug.update(dict(O=O, sdf=sdf, nrm=nrm, alpha_fg=alpha_fg, alpha_bg=alpha_bg, w=w, T=T, pix=pix,
                                   nhat=nhat, RAW=RAW, ATT=ATT, attpix=attpix, dirs_u=dirs_u, dirs_s=dirs_s,
                                   elraw=elraw, svraw=svraw, color=color, colbg=colbg, bgraw=bgraw, x_fg=x_fg,
                                   t_fg=t_fg, x_bg=x_bg, t_bg=t_bg, mask=mask, dims=(B, R, N, Nb, M)))
        if not backward:
            self._weights_synced = False
            return losses[0, :N_LOSSES]

        # ======================================= backward =======================================
        d_el = self.buf("d_el", rows_d, 4)
        d_sv = self.buf("d_sv", rows_d, 4)
        d_attpix = self.buf("d_attpix", NR, 12)
        d_nhat = self.buf("d_nhat", NR, 3)
        d_colbg = self.buf("d_colbg", NR, 3)
        self.call("ndjir_shade_backward", NR, M, P_(nhat), P_(attpix), P_(raydir), P_(dirs_u), P_(dirs_s), P_(elraw), 4,
                  P_(svraw), 4, P_(colbg), P_(color_gt), cfg5, P_(d_el), P_(d_sv), P_(d_attpix), P_(d_nhat),
                  P_(d_colbg))
        # environment light: parameters only (directions carry no gradient, sampler.py:391)
        self.mlp_backward("el", "el", P_(Xel), ldel, rows_d, acts["el"], [(P_(d_el), 4)])
        # soft visibility: input gradient -> per-ray sums
        dXsv = self.buf("dXsv", rows_d, r4(Df + 6))
        self.mlp_backward("sv", "sv", P_(Xsv), ldsv, rows_d, acts["sv"], [(P_(d_sv), 4)], dX=P_(dXsv),
                          lddx=dXsv.shape[1], dx_cols=Df + 6)
        dpix = self.buf("dpix", NR, LDO, zero=True)
        self.call("ndjir_group_sum", NR, M, Df + 6, P_(dpix), LDO, P_(dXsv), dXsv.shape[1], 0)
        self.call("ndjir_group_sum", NR, M, Df + 6, P_(dpix), LDO, P_(dXsv, NR * M * dXsv.shape[1]), dXsv.shape[1], 1)
        # dpix columns Df+3:Df+6 currently hold d nhat from the visibility input; add the shading part, then
        # turn d nhat into d n_pix
        self.copy2d(NR, 3, P_(d_nhat), 3, P_(dpix, Df + 3), LDO, accum=1)
        self.call("ndjir_pixel_normal_backward", NR, P_(pix, Df + 3), LDO, float(r.eps_normal), P_(d_nhat),
                  P_(dpix, Df + 3), LDO, 0)
        # weights gradient
        dw = self.buf("dw", NR, S, zero=True)
        dO = self.buf("dO", P, LDO)
        self.call("ndjir_volume_render_backward", NR, N, Df + 6, P_(w), S, P_(O), LDO, P_(dpix), LDO, P_(dO), LDO, 0,
                  P_(dw), S)
        dATT = self.buf("dATT", P, 12)
        self.call("ndjir_volume_render_backward", NR, N, 12, P_(w), S, P_(ATT), 12, P_(d_attpix), 12, P_(dATT), 12, 0,
                  P_(dw), S)
        d_bgraw = self.buf("d_bgraw", rows_bg, 4)
        self.call("ndjir_bg_color_backward", NR, Nb, P_(w, N), S, P_(bgraw), 4, P_(d_colbg), P_(dw, N), S,
                  P_(d_bgraw), 4)
        # material heads
        dRAW = self.buf("dRAW", P, 16)
        self.call("ndjir_sample_attributes_backward", P, N, P_(RAW), P_(dATT), P_(nrm), 3, P_(maskv), cfg10, inv_denorm,
                  P_(dRAW), P_(dO, Df + 3), LDO)
        self.mlp_backward("bc", "bc", P_(O), LDO, P, acts["bc"], [(P_(dRAW, 0), 16)], dX=P_(dO), lddx=LDO,
                          accum_dx=True, dx_cols=Df + 3)
        for name, col in (("ii", 3), ("ro", 4), ("sp", 6)):
            self.mlp_backward(name, name, P_(O), LDO, P, acts[name], [(P_(dRAW, col), 16)], dX=P_(dO), lddx=LDO,
                              accum_dx=True, dx_cols=Df + 6)
        dXpl = self.buf("dXpl", P, r4(Df + 6))
        self.mlp_backward("pl", "pl", P_(Xpl), ldpl, P, acts["pl"], [(P_(dRAW, 12), 16)], dX=P_(dXpl),
                          lddx=dXpl.shape[1], dx_cols=Df + 6)
        self.copy2d(P, Df + 6, P_(dO), LDO, P_(dXpl), dXpl.shape[1], accum=1)
        # compositing + alpha
        dalpha_fg = self.buf("dalpha_fg", P, 1)
        dalpha_bg = self.buf("dalpha_bg", rows_bg, 1)
        self.call("ndjir_composite_backward", NR, N, Nb, P_(alpha_fg), P_(maskv), P_(alpha_bg), P_(T), P_(dw),
                  P_(dalpha_fg), P_(dalpha_bg))
        dsdf = self.buf("dsdf", P, 1, zero=True)
        g_gain = ps.grad.data_ptr() + 4 * ps.gain_off
        self.call("ndjir_neus_alpha_backward", P, N, P_(dalpha_fg), P_(sdf), P_(nrm), 3, P_(raydir), P_(t_fg), gain_p,
                  float(cos_anneal_ratio), P_(dsdf), P_(dO, Df + 3), LDO, g_gain)
        # background networks
        dXbg1 = self.buf("dXbg1", rows_bg, r4(Dfb))
        self.mlp_backward("bg1", "bg1", P_(Xbg1), ldb1, rows_bg, acts_bg1, [(P_(d_bgraw), 4)], dX=P_(dXbg1),
                          lddx=dXbg1.shape[1], dx_cols=Dfb)
        d_dens = self.buf("d_dens", rows_bg, 1)
        self.call("ndjir_bg_alpha_backward", rows_bg, Nb, P_(dalpha_bg), P_(dens), 1, P_(t_bg), P_(d_dens), 1)
        self.mlp_backward("bg0", "bg0", P_(Xbg0), Xbg0.shape[1], rows_bg, acts_bg0,
                          [(P_(d_dens), 1), (P_(dXbg1), dXbg1.shape[1])])
        # geometric network: second-order terms from d L / d normal, then the standard sweep
        nbar = self.buf("nbar", P, 3)
        self.copy2d(P, 3, P_(nbar), 3, P_(dO, Df + 3), LDO)
        Z2 = self.geo_normal_adjoint(x_fg, P, "main", A, GZ, Gin, nbar)
        self.geo_backward(x_fg, P, "main", A, dsdf, dO, Z2)
        # perturbed branch
        dOp = self.buf("dO_ptb", P, LDO)
        self.mlp_backward("bc", "bcp", P_(Op), LDO, P, acts["bcp"], [(P_(dRAW, 13), 16)], dX=P_(dOp), lddx=LDO,
                          dx_cols=Df + 3)
        self.geo_backward(x_ptb, P, "ptb", Ap, None, dOp, None)
        # TV
        if tv_on:
            for part, width, off in self._grid_parts():
                tvg = self.buf(f"tvg_{part}", P, width)
                self.call("ndjir_ray_mask_fill", P, N, width, P_(tvg), P_(maskv), inv_denorm, float(tr.tv_weight))
                self._grid_scatter("tv_bwd", part, P, x_fg, [tvg], with_feature=True)
        if self.world_size > 1:
            allreduce_gradients(ps, self.pg, include_grid=not self._sparse_grid())
        if keep:
            self.debug.update(dict(dO=dO, dsdf=dsdf, dw=dw, dRAW=dRAW, d_attpix=d_attpix, dpix=dpix, nbar=nbar,
                                   dalpha_fg=dalpha_fg, dalpha_bg=dalpha_bg, d_el=d_el, d_sv=d_sv))
        self._weights_synced = False
        return losses[0, :N_LOSSES]
