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
import ctypes
import math

import numpy as np
import torch

from . import _lib
from . import h16
from .h16 import HMat
from .config import check_supported, grid_channels
from .parallel import allgather_rows, allreduce_gradients, allreduce_mask_sum
from .scene import network_dims, pe_dim, NET_ORDER

EPI_BIAS, EPI_SOFTPLUS, EPI_ACCUM, EPI_MUL_S, EPI_ADJ, EPI_ATOMIC = range(6)
LOSS_NAMES = ["loss", "loss_rgb", "loss_eikonal", "loss_tv", "loss_mask", "prior_base_color", "prior_roughness",
              "prior_specular_reflectance", "reg_std_roughness", "reg_std_specular_reflectance"]
N_LOSSES = len(LOSS_NAMES)


def r4(n):
    return (n + 3) // 4 * 4


def r8(n):
    return (n + 7) // 8 * 8


class Layer:
    """One affine layer in INTERNAL layout: W (K, ldw) row-major with ldw = round4(N); `rowmap[i]` is the internal
    row of reference input row i (the engine keeps head inputs as [feature | x | normal | ...] so the geometric
    feature block is 16-byte aligned and shared between heads; the reference order is [x | ... | feature | normal])."""

    def __init__(self, K, N, rowmap=None, col0=0, Nref=None):
        self.K, self.N = K, N
        self.ldw = r8(N)
        self.rowmap = np.arange(K) if rowmap is None else np.asarray(rowmap)
        self.col0 = col0            # first reference output column held by this layer (split last layers)
        self.w_off = self.b_off = self.t_off = -1
        self.ldt = r8(K)            # row stride of the transposed copy W^T (N, ldt)


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
        inv = [("inv", 1)] if pl.use_inverse_distance else []          # network.py:410 (config/no_inv_distance_square.yaml)
        self.nets["pl"] = head("pl", [("x", 3), ("pe", pe_dim(3, pl.pe_bands)), ("f", Df), ("n", 3)] + inv,
                               ["f", "x", "n", "pe"] + [n for n, _ in inv])
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
                L.b_off = n; n += r8(L.N)
        self.gain_off = n; n += 8
        self.n_mlp = n
        nt = 0
        for name in NET_ORDER:
            for L in self.nets[name]:
                L.t_off = nt; nt += r8(L.N) * L.ldt
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
                    nt += r8(r4(L.N) * r4(rows))
        # transposed weight copies for the input-gradient products (refreshed once per step, engine.refresh_transposes)
        self.data_t = torch.zeros(max(r8(nt), 8), dtype=torch.float32, device=device)
        self.data = torch.zeros(n, dtype=torch.float32, device=device)
        # lo parts (x - tf32(x)) of both weight buffers for the tensor-core products (registered with the library in
        # Engine.__init__, refreshed together with the transposes)
        self.data_lo = torch.zeros_like(self.data)
        self.data_t_lo = torch.zeros_like(self.data_t)
        self.grad = torch.zeros(n, dtype=torch.float32, device=device)
        # split-fp16 copies of both weight buffers (same element offsets; planes [2, n]) and their common power-of-two
        # scale: the B operands of the split-fp16 engine (W^T for forward products, W for input-gradient products),
        # refreshed with the transposes
        self.w16 = torch.zeros((2, n), dtype=torch.float16, device=device)
        self.wt16 = torch.zeros((2, self.data_t.numel()), dtype=torch.float16, device=device)
        self.wscale = torch.ones(4, dtype=torch.float32, device=device)      # [scale, amax, -, -]
        self.wflags = torch.zeros(4, dtype=torch.int32, device=device)
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

    def reference_dict(self):
        """Current parameters in the layout scene.init_params / load_reference use."""
        ex = self.export_reference("data")
        P = {net: [(ex[f"{net}.W{i}"], ex[f"{net}.b{i}"]) for i, _ in self._ref_layers(net)] for net in NET_ORDER}
        P["geo_gain"] = ex["geo_gain"]
        P["pl_gain"] = np.asarray([self.pl_gain], np.float32)
        P["grid"] = {k: ex[f"grid.{k}"] for k in self.grid}
        return P

    def save_parameters(self, path):
        """`nn.save_parameters(path)` of the reference (python/train.py:101): an nnabla `.h5` parameter file with the
        reference's scope names, (in, out) weight layout, `need_grad` / `index` attributes (ndjir_b200/h5lite.py)."""
        from . import h5lite, nnabla_names
        params = nnabla_names.to_nnabla(self.conf, self.reference_dict())
        # the photogrammetric-light gain is scheduled, not trained (network.py:418-420)
        ng = {n: not n.endswith("photogrammetric-light-network/gain") for n in params}
        h5lite.save_parameters(path, params, need_grad=ng)

    def load_parameters(self, path):
        """`nn.load_parameters(path)` of the reference (python/render_image.py:43, extract_by_mc.py:300)."""
        from . import h5lite, nnabla_names
        params, _ = h5lite.load_parameters(path)
        self.load_reference(nnabla_names.from_nnabla(self.conf, params))

    def zero_grad(self, stream=None):
        """zero every gradient buffer with the library's own fill kernel (ndjir_fill) on `stream`"""
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.call("ndjir_fill", self.grad.numel(), self.grad, 0.0, st)
        for v in self.grid_grad.values():
            _lib.call("ndjir_fill", v.numel(), v, 0.0, st)

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

    def W16(self, L, row0=0):
        """rows [row0, ...) of W_L as split-fp16 planes (K_in rows x N): the K-major B operand of input-gradient products"""
        off = 2 * (L.w_off + row0 * L.ldw)
        return HMat(self.w16.data_ptr() + off, self.w16.data_ptr() + 2 * self.w16.shape[1] + off, L.ldw,
                    self.wscale.data_ptr(), None)

    def WT16(self, L):
        """W_L^T as split-fp16 planes (N rows x K_in): the K-major B operand of forward products"""
        off = 2 * L.t_off
        return HMat(self.wt16.data_ptr() + off, self.wt16.data_ptr() + 2 * self.wt16.shape[1] + off, L.ldt,
                    self.wscale.data_ptr(), None)

    def Bt32(self, L, row0=0):
        """fp32 operand B(k, j) = W[row0 + j, k] of the memory-bound input-gradient products as (pointer, b_rs, b_cs)"""
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


class Mat:
    """A matrix the MLP passes read or write: fp32 rows (`f`, a (rows_alloc, ldf) tensor) and / or split-fp16 planes
    (`h`, an h16.HBuf).  `rs` overrides the fp32 row stride (0 = one row broadcast to every sample)."""
    __slots__ = ("f", "h", "cols", "rows", "rs")

    def __init__(self, f=None, h=None, cols=0, rows=None, rs=None):
        self.f, self.h, self.cols, self.rs = f, h, cols, rs
        self.rows = rows if rows is not None else (f.shape[0] if f is not None else (h.rows if h is not None else 0))

    @property
    def ldf(self):
        return self.rs if self.rs is not None else self.f.shape[1]

    def fptr(self, col=0, row=0):
        return self.f.data_ptr() + 4 * (row * self.ldf + col)

    def hmat(self, col=0, track=True, row=0):
        return self.h.hmat(col, row=row, track=track)


class Engine:
    def __init__(self, conf, device="cuda", world_size=1, process_group=None, grid_exchange="auto", mlp="h16"):
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
        # a configuration that takes a branch the kernels do not have must fail here instead of silently rendering
        # default.yaml's (config.check_supported names the key)
        check_supported(conf)
        g = conf.geometric_network
        self.Df = g.feature_size
        self.Dg = grid_channels(conf)
        self.din = pe_dim(3, g.pe_bands) + self.Dg
        self.npe = pe_dim(3, g.pe_bands)
        self.ld0 = r4(self.din)
        self.LDO = r4(self.Df + 6)
        self.skip = g.skip_layers[0] if len(g.skip_layers) else -1
        ii = conf.implicit_illumination_network
        # networks the configuration switches off (config/no_implicit_illumination.yaml, no_lightp.yaml): their parameters
        # stay in the store (zero gradient, left out of parameter files like the reference, scene.active_nets)
        self.use_ii, self.use_pl = bool(ii.use_me), bool(conf.photogrammetric_light_network.use_me)
        self.cskip = 1.0 / math.sqrt(2.0) if g.use_inv_square else 1.0
        self._bufs = {}
        self._graphs = {}
        self._reserve = 0
        self.one = torch.ones(4, dtype=torch.float32, device=self.device)
        self.ones_col = Mat(f=self.one.view(1, 4), rs=0)     # a column of ones (rank-1 products with w_sdf)
        # MLP engine: "h16" = split-fp16 storage + tcgen05 kind::f16 products (csrc/gemm_h.cu); "fp32" = fp32 storage,
        # 3xTF32 tcgen05 products or, with ndjir_set_option("mlp_tensor_cores", 0), the exact FFMA parity path
        assert mlp in ("h16", "fp32")
        self.h16 = mlp == "h16"
        self.scales = h16.Scales(self.device) if self.h16 else None
        self._calibrated = not self.h16
        self.precise_fwd = True
        self.fused_sampler = True     # sample_points as one C-ABI call (ndjir_sample_points_fwd); False: sequenced here
        self.rad = float(conf.renderer.bounding_sphere_radius)
        self.debug = {}
        self.profile, self.prof_events, self.n_launches = False, [], 0

    # ------------------------------------------------------------------------------------------------
    def stream(self):
        return torch.cuda.current_stream().cuda_stream

    @property
    def fused_calls(self):
        """stages as single C-ABI calls (csrc/fused_path.cu).  The instrumented step (profile: CUDA events around every
        product) sequences the same products call by call instead, so that each one can be timed."""
        return self.h16 and self.fused_sampler and not self.profile

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
            if dtype == torch.float32:
                self.call("ndjir_fill", rows * cols, t, 0.0)
            else:
                t[:rows].zero_()
        return t

    # ------------------------------------------------------------------------------------------------
    # matrices of the MLP passes: fp32 rows and / or split-fp16 planes (class Mat), products in either engine
    # ------------------------------------------------------------------------------------------------
    def mat(self, name, rows, cols, kind="a", zero=False, slot=None, grad=False):
        """kind 'f': fp32 rows (row stride round4(cols)); 'a': the MLP activation format (split-fp16 planes in h16 mode,
        fp32 rows otherwise); 'fa': fp32 rows plus, in h16 mode, planes that sync_h() refreshes from them.
        grad=True marks a gradient-magnitude tensor (initial power-of-two scale 2^24 instead of 2^4)."""
        key = ("mat", name, cols, kind)
        m = self._bufs.get(key)
        if m is None or m.rows < rows:
            need = max(rows, self._reserve)
            f = h = None
            if kind in ("f", "fa") or not self.h16:
                f = torch.empty((need, r4(cols)), dtype=torch.float32, device=self.device)
            if self.h16 and kind in ("a", "fa"):
                h = h16.HBuf(need, cols, self.device, self.scales, slot or name, init=2.0 ** 24 if grad else 16.0)
            m = Mat(f, h, cols, need)
            self._bufs[key] = m
        if zero and m.f is not None:      # (planes are always rewritten by sync_h / an epilogue before they are read)
            self.call("ndjir_fill", rows * m.f.shape[1], m.f, 0.0)
        return m

    def sync_h(self, m, cols, rows, col=0):
        """planes of a 'fa' matrix <- its fp32 rows (columns [col, col + cols)); nothing to do in fp32 mode"""
        if self.h16 and m.h is not None:
            self._pack(rows, cols, m.fptr(col), m.ldf, 1, 1.0, m, col, 0)

    def _pack(self, rows, ncols, src, ld_src, rep, alpha, dst, dcol, drow):
        """fp32 columns -> planes (ndjir_pack_h: 16-byte groups of 8 columns, a ragged last group inside the same launch)"""
        self.call("ndjir_pack_h", rows, ncols, src, ld_src, rep, alpha, dst.hmat(dcol, row=drow))

    def fill_cols(self, dst, dcol, src, ld_src, ncols, rows, rep=1, alpha=1.0, drow=0):
        """dst[drow + r, dcol:dcol+ncols] = alpha * src[r // rep, :ncols]   (src: fp32 device address)"""
        if dst.f is not None:
            self.copy2d(rows, ncols, dst.fptr(dcol, drow), dst.ldf, src, ld_src, rep=rep, alpha=alpha)
        else:
            self._pack(rows, ncols, src, ld_src, rep, alpha, dst, dcol, drow)

    def copy_cols(self, dst, src, ncols, rows):
        """dst[:, :ncols] = src[:, :ncols] between two matrices of the activation format (same scale slot in h16 mode)"""
        if self.h16:
            n8 = r8(ncols) if (r8(ncols) <= dst.h.ld and r8(ncols) <= src.h.ld) else ncols
            self.call("ndjir_copy2d_h", rows, n8, dst.hmat(0, track=False), src.hmat(0, track=False), 1)
        else:
            self.copy2d(rows, ncols, dst.fptr(), dst.ldf, src.fptr(), src.ldf)

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

    def gemm_h(self, M, N, K, epi, **kw):
        """one product of the split-fp16 engine (ndjir_gemm_h)"""
        if self.profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.n_launches += 1
        h16.gemm_h(self.stream(), M, N, K, epi, **kw)
        if self.profile:
            e1.record()
            mn = kw.get("mn_major", False)
            self.prof_events.append((e0, e1, 2.0 * M * N * K, f"epi{epi} {M}x{N}x{K}" + (" mn" if mn else "")))

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

    @staticmethod
    def _out(Y, col, key="C"):
        """keyword arguments of ndjir_gemm_h for an output / epilogue operand in whichever form the matrix has"""
        if Y is None:
            return {}
        if key in ("C", "C2"):
            if Y.f is not None:
                return {key: Y.fptr(col), "ld" + key.lower(): Y.ldf}
            return {key + "h": Y.hmat(col)}
        if Y.h is not None:
            return {key + "h": Y.hmat(col, track=False)}
        return {key: Y.fptr(col), "ld" + key.lower(): Y.ldf}

    def fwd(self, rows, L, X, Y, ycol, epi, out_scale=1.0):
        """Y[:, ycol:ycol+N] = epi(X[:, :K] W_L + b_L)   (PF.affine [+ softplus], python/network.py:84-93)"""
        ps = self.params
        if not self.h16:
            self.gemm(rows, L.N, L.K, X.fptr(), X.ldf, 1, ps.W(L), L.ldw, 1, Y.fptr(ycol), Y.ldf, epi, bias=ps.b(L),
                      out_scale=out_scale)
        elif L.N <= 8:     # one pass over the activations (sdf, colours, ...): fp32 weights, fp32 output
            self.gemm_h(rows, L.N, L.K, epi, A=X.hmat(0, track=False), B32=ps.W(L), b_rs=L.ldw, b_cs=1, bias=ps.b(L),
                        C=Y.fptr(ycol), ldc=Y.ldf)
        else:
            self.gemm_h(rows, L.N, L.K, epi, A=X.hmat(0, track=False), B=ps.WT16(L), precise=self.precise_fwd,
                        bias=ps.b(L), out_scale=out_scale, **self._out(Y, ycol))

    def dgrad(self, rows, L, dY, dycol, dX, dxcol, epi, row0=0, ncols=None, alpha=1.0, H=None, hscale=1.0, U=None,
              precise=False):
        """dX[:, dxcol:dxcol+ncols] = epi(alpha * dY[:, dycol:dycol+N] W_L[row0:row0+ncols, :]^T)"""
        ps = self.params
        ncols = L.K if ncols is None else ncols
        if not self.h16:
            self.gemm(rows, ncols, L.N, dY.fptr(dycol), dY.ldf, 1, *ps.Bt(L, row0), dX.fptr(dxcol), dX.ldf, epi,
                      alpha=alpha, H=(H.fptr() if H is not None else 0), ldh=(H.ldf if H is not None else 0),
                      hscale=hscale, U=(U.fptr() if U is not None else 0), ldu=(U.ldf if U is not None else 0))
            return
        kw = dict(alpha=alpha, hscale=hscale, **self._out(dX, dxcol), **self._out(H, 0, "H"), **self._out(U, 0, "U"))
        b32, b_rs, b_cs = ps.Bt32(L, row0)
        if dY.h is None:       # narrow fp32 gradient (<= 8 columns): rank-N update with the fused epilogue
            self.gemm_h(rows, ncols, L.N, epi, A32=dY.fptr(dycol), a_rs=dY.ldf, a_cs=1, B32=b32, b_rs=b_rs, b_cs=b_cs,
                        **kw)
        elif ncols <= 8:       # narrow fp32 result (grid-feature gradient)
            self.gemm_h(rows, ncols, L.N, epi, A=dY.hmat(dycol, track=False), B32=b32, b_rs=b_rs, b_cs=b_cs, **kw)
        elif 256 < ncols <= 264 and dX.f is not None and epi in (EPI_BIAS, EPI_ACCUM) and H is None and U is None:
            # [feature(256) | x | normal] input gradients: one 256-wide tensor-core tile + a memory-bound pass for the few
            # remaining columns (a second 256-wide tile would stream dY again for 3-6 outputs)
            self.dgrad(rows, L, dY, dycol, dX, dxcol, epi, row0=row0, ncols=256, alpha=alpha, precise=precise)
            self.dgrad(rows, L, dY, dycol, dX, dxcol + 256, epi, row0=row0 + 256, ncols=ncols - 256, alpha=alpha)
        else:
            self.gemm_h(rows, ncols, L.N, epi, A=dY.hmat(dycol, track=False), B=ps.W16(L, row0), precise=precise,
                        **kw)

    def wgrad_(self, rows, L, X, dY, dycol, kin=None, bias=True):
        """gW_L[:kin, :] += X[:, :kin]^T dY[:, dycol:dycol+N];  gb_L += column sums of dY"""
        ps = self.params
        kin = L.K if kin is None else kin
        if not self.h16:
            self.wgrad(rows, kin, L.N, X.fptr(), X.ldf, dY.fptr(dycol), dY.ldf, ps.gW(L), L.ldw,
                       gb=(ps.gb(L) if bias else 0))
            return
        if dY.h is None:
            self.gemm_h(kin, L.N, rows, EPI_ATOMIC, A=X.hmat(0, track=False), mn_major=True, B32=dY.fptr(dycol),
                        b_rs=dY.ldf, b_cs=1, C=ps.gW(L), ldc=L.ldw)
            if bias:
                self.call("ndjir_colsum", rows, L.N, ps.gb(L), dY.fptr(dycol), dY.ldf, 1.0)
            return
        tiles = ((kin + 127) // 128) * ((L.N + 255) // 256)
        split = max(1, min(rows // 256, (148 * 2) // tiles))
        self.gemm_h(kin, L.N, rows, EPI_ATOMIC, A=X.hmat(0, track=False), B=dY.hmat(dycol, track=False), mn_major=True,
                    split_k=split, C=ps.gW(L), ldc=L.ldw, colsum=(ps.gb(L) if bias else None))

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
        """W^T copies used by the input-gradient products (fp32 engine) / forward products (split-fp16 engine), and the
        low-order / split-fp16 copies of the weights; call after the parameters change (once per step)."""
        ps = self.params
        for name in NET_ORDER:
            for L in ps.nets[name]:
                self.call("ndjir_transpose", L.K, L.N, ps.data_t.data_ptr() + 4 * L.t_off, L.ldt, ps.W(L), L.ldw)
                sk = ps.skip_t.get(id(L))
                if sk is not None and not self.h16:   # rows [row0, row0 + rows) of W -> their own transposed block
                    self.call("ndjir_transpose", sk[1], L.N, ps.data_t.data_ptr() + 4 * sk[2], sk[3], ps.W(L, sk[0]), L.ldw)
        if self.h16:
            # one power-of-two scale for all weights from their current maximum, then both buffers as fp16 planes
            n, nt = ps.data.numel(), ps.data_t.numel()
            self.call("ndjir_amax", n, ps.data, P_(ps.wscale, 1))
            self.call("ndjir_scale_update", 1, ps.wscale, P_(ps.wscale, 1), ps.wflags, self.scales.target_log2)
            sc = ps.wscale.data_ptr()
            self.call("ndjir_pack_h", 1, n, ps.data, n, 1, 1.0, HMat(ps.w16.data_ptr(), ps.w16.data_ptr() + 2 * n, n, sc, None))
            self.call("ndjir_pack_h", 1, nt, ps.data_t, nt, 1, 1.0,
                      HMat(ps.wt16.data_ptr(), ps.wt16.data_ptr() + 2 * nt, nt, sc, None))
            return
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
        """A0 (rows, ld0) fp32 = [PE(x) | grid features]"""
        g = self.conf.geometric_network
        self.call("ndjir_positional_encoding", rows, 3, g.pe_bands, P_(x), 3, 1, P_(A0), self.ld0)
        for part, width, off in self._grid_parts():
            tmp = self.buf(f"gq_{part}", rows, width)
            self._grid_call("query", part, rows, P_(tmp), P_(x), P_(self.params.grid[part]))
            self.copy2d(rows, width, P_(A0, self.npe + off), self.ld0, P_(tmp), width)
        if self.ld0 > self.din:
            self.call("ndjir_copy2d", rows, self.ld0 - self.din, P_(A0, self.din), self.ld0, P_(self.one), 0, 1, 0.0, 0)

    def geo_forward(self, x, rows, tag, store, want_feat=True, sdf_out=None, O=None):
        """Forward pass.  store=True keeps every layer input in matrices `{tag}_A{l}` for the backward passes.
        O: 'fa' matrix whose fp32 columns [0, Df) receive the feature.  Returns (A list, sdf tensor)."""
        net = self.params.nets["geo"]
        nl = len(net) - 2          # hidden layers (reference layers 0..L-2)
        A0 = self.mat(f"{tag}_A0", rows, self.din, "fa")
        if self.fused_calls and store:
            # ONE C-ABI call (ndjir_geo_forward, csrc/fused_path.cu) on the same buffers: encoding, grid query, every layer
            # with its input kept, sdf column, feature block
            A = [A0] + [self.mat(f"{tag}_A{l + 1}", rows, net[l + 1].K, "a") for l in range(nl)]
            sdf = sdf_out if sdf_out is not None else self.buf(f"{tag}_sdf", rows, 1)
            gw = sum(w for _, w, _ in self._grid_parts())
            ws = h16.GeoStore()
            ws.enc, ws.ld_enc = A0.f.data_ptr(), self.ld0
            ws.grid_tmp = self.buf("gq_fused", rows, max(gw, 1)).data_ptr() if gw else None
            for l, a in enumerate(A):
                ws.acts[l] = a.hmat(0)
            feat = (O.fptr(0), O.ldf) if want_feat else (None, 0)
            self.n_launches += nl + 6
            self.call("ndjir_geo_forward", self.geo_net_desc(), rows, P_(x), P_(sdf), feat[0], feat[1], ws)
            return A, sdf
        self.geo_input(x, rows, A0.f, tag)
        self.sync_h(A0, self.din, rows)
        A = [A0]
        for l in range(nl):
            L = net[l]
            nxt = self.mat(f"{tag}_A{l + 1}" if store else f"geo_pp{l % 2}", rows, net[l + 1].K, "a")
            self.fwd(rows, L, A[l], nxt, 0, EPI_SOFTPLUS, out_scale=self.cskip if (l + 1) == self.skip else 1.0)
            if (l + 1) == self.skip:
                self.fill_cols(nxt, L.N, A0.fptr(), A0.ldf, self.din, rows, alpha=self.cskip)
            A.append(nxt)
        Ls, Lf = net[-2], net[-1]
        sdf = sdf_out if sdf_out is not None else self.buf(f"{tag}_sdf", rows, 1)
        self.fwd(rows, Ls, A[-1], Mat(f=sdf), 0, EPI_BIAS)
        if want_feat:
            self.fwd(rows, Lf, A[-1], O, 0, EPI_BIAS)
        return A, sdf

    def geo_normal(self, x, rows, tag, A, nrm):
        """n = d sdf / d x by a reverse sweep (the nn.grad graph of renderer.py:52).  Keeps GZ[l] = dsdf/dz_l."""
        net = self.params.nets["geo"]
        nl = len(net) - 2
        Ls = net[-2]
        GZ = [self.mat(f"{tag}_GZ{l}", rows, net[l].N, "a") for l in range(nl)]
        if self.fused_calls:
            # ONE C-ABI call (ndjir_geo_normal): the reverse sweep, the encoding's and the grid's input gradients
            Gin = self.mat(f"{tag}_Gin", rows, self.din, "f")
            gw = sum(w for _, w, _ in self._grid_parts())
            fw = h16.GeoStore()
            fw.enc, fw.ld_enc, fw.grid_tmp = A[0].f.data_ptr(), self.ld0, None
            for l, a in enumerate(A):
                fw.acts[l] = a.hmat(0)
            ws = h16.GeoNormalWs()
            for l, gz in enumerate(GZ):
                ws.gz[l] = gz.hmat(0)
            ws.g_in = Gin.f.data_ptr()
            ws.grid_tmp = self.buf("gg_fused", rows, max(gw, 1)).data_ptr() if gw else None
            ws.ones = self.one.data_ptr()
            self.n_launches += nl + 5
            self.call("ndjir_geo_normal", self.geo_net_desc(), rows, P_(x), fw, ws, P_(nrm), 3)
            return GZ, Gin
        Gin = self.mat(f"{tag}_Gin", rows, self.din, "f", zero=True)
        c = self.cskip
        # top: gA_{L-1}[p,:] = w_sdf ; GZ[nl-1] = gA * s
        self.dgrad(rows, Ls, self.ones_col, 0, GZ[nl - 1], 0, EPI_MUL_S, H=A[nl])
        skip_wrote = False
        for l in range(nl - 1, 0, -1):
            L = net[l]
            is_skip = (l == self.skip)
            n_prev = net[l - 1].N
            self.dgrad(rows, L, GZ[l], 0, GZ[l - 1], 0, EPI_MUL_S, ncols=n_prev, alpha=(c if is_skip else 1.0),
                       H=A[l], hscale=(1.0 / c if is_skip else 1.0), precise=self.precise_fwd)
            if is_skip:
                self.dgrad(rows, L, GZ[l], 0, Gin, 0, EPI_BIAS, row0=n_prev, ncols=self.din, alpha=c,
                           precise=self.precise_fwd)
                skip_wrote = True
        self.dgrad(rows, net[0], GZ[0], 0, Gin, 0, EPI_ACCUM if skip_wrote else EPI_BIAS, ncols=self.din,
                   precise=self.precise_fwd)
        g = self.conf.geometric_network
        self.call("ndjir_positional_encoding_grad_input", rows, 3, g.pe_bands, A[0].fptr(), self.ld0, Gin.fptr(),
                  self.ld0, P_(nrm), 3, 0)
        for part, width, off in ([] if g.voxel.use_ste else self._grid_parts()):     # (use_ste: voxel_feature.py:390-391)
            tmp = self.buf(f"gg_{part}", rows, width)
            self.copy2d(rows, width, P_(tmp), width, Gin.fptr(self.npe + off), self.ld0)
            self._grid_call("grad_query", part, rows, P_(nrm), P_(tmp), P_(x), P_(self.params.grid[part]))
        return GZ, Gin

    def geo_normal_adjoint(self, x, rows, tag, A, GZ, Gin, nbar):
        """Reverse of geo_normal given nbar = dL/dn (P,3): accumulates the weight gradients of the normal pass,
        the grid gradient through d/dq, and returns Z2[l] = second-order contribution to dL/dz_l."""
        ps, net = self.params, self.params.nets["geo"]
        nl = len(net) - 2
        g = self.conf.geometric_network
        c = self.cskip
        Gh0 = self.mat(f"{tag}_Gh0", rows, self.din, "fa", zero=True, grad=True)
        self.call("ndjir_positional_encoding_grad_input_adjoint", rows, 3, g.pe_bands, A[0].fptr(), self.ld0, P_(nbar), 3,
                  Gh0.fptr(), self.ld0)
        for part, width, off in ([] if g.voxel.use_ste else self._grid_parts()):     # (use_ste: no grid terms in the normal)
            tmp = self.buf(f"gg_{part}", rows, width)       # g_in grid columns (dense)
            self.copy2d(rows, width, P_(tmp), width, Gin.fptr(self.npe + off), self.ld0)
            tmp2 = self.buf(f"ggo_{part}", rows, width)
            self._grid_call("ggo", part, rows, P_(tmp2), P_(nbar), P_(x), P_(self.params.grid[part]))
            self.copy2d(rows, width, Gh0.fptr(self.npe + off), self.ld0, P_(tmp2), width)
            self._grid_scatter("gqgf", part, rows, x, [nbar, tmp])
        self.sync_h(Gh0, self.din, rows)
        if self.fused_calls:
            # the upward walk as ONE C-ABI call (ndjir_geo_normal_adjoint, csrc/fused_path.cu)
            Gh = [self.mat(f"{tag}_Gh{l + 1}", rows, net[l + 1].K, "a", grad=True) for l in range(nl)]
            Z2 = [self.mat(f"{tag}_Z2{l}", rows, net[l].N, "a", grad=True) for l in range(nl)]
            st = h16.GeoStore()
            for l in range(nl + 1):
                st.acts[l] = A[l].hmat(0, track=False)
            arr = lambda ms, track: (h16.HMat * nl)(*[m.hmat(0, track=track) for m in ms])
            self.n_launches += 2 * nl + (1 if 0 < self.skip <= nl else 0)
            self.call("ndjir_geo_normal_adjoint", self.geo_net_desc(), self._grads(net[:nl]), self._grads([net[-2]]), rows,
                      st, arr(GZ, False), Gh0.fptr(), Gh0.ldf, Gh0.hmat(0, track=False), arr(Gh, True), arr(Z2, True),
                      self.ones_col.fptr())
            return Z2
        Ghat, Z2 = [Gh0], []
        for l in range(nl):
            L = net[l]
            nxt = self.mat(f"{tag}_Gh{l + 1}", rows, net[l + 1].K, "a", grad=True)
            z2 = self.mat(f"{tag}_Z2{l}", rows, L.N, "a", grad=True)
            is_skip = (l + 1) == self.skip
            Gh = Ghat[l]
            osc, hsc = (c if is_skip else 1.0), (1.0 / c if is_skip else 1.0)
            if self.h16:
                self.gemm_h(rows, L.N, L.K, EPI_ADJ, A=Gh.hmat(0, track=False), B=ps.WT16(L), out_scale=osc, hscale=hsc,
                            Hh=A[l + 1].hmat(0, track=False), Uh=GZ[l].hmat(0, track=False), Ch=z2.hmat(),
                            C2h=nxt.hmat())
            else:
                self.gemm(rows, L.N, L.K, Gh.fptr(), Gh.ldf, 1, ps.W(L), L.ldw, 1, z2.fptr(), z2.ldf, EPI_ADJ,
                          out_scale=osc, H=A[l + 1].fptr(), ldh=A[l + 1].ldf, hscale=hsc, U=GZ[l].fptr(),
                          ldu=GZ[l].ldf, C2=nxt.fptr(), ldc2=nxt.ldf)
            if is_skip:
                self.fill_cols(nxt, L.N, Gh0.fptr(), Gh0.ldf, self.din, rows, alpha=c)
            self.wgrad_(rows, L, Gh, GZ[l], 0, bias=False)
            Ghat.append(nxt)
            Z2.append(z2)
        Ls = net[-2]
        # g w_sdf[k] += sum_p Ghat_top[p,k]
        self.wgrad_(rows, Ls, Ghat[nl], self.ones_col, 0, bias=False)
        return Z2

    def geo_backward(self, x, rows, tag, A, dsdf, dO, Z2):
        """Standard reverse sweep.  dsdf (rows,1) tensor or None, dO: 'fa' matrix holding dL/dfeature in its fp32
        columns 0:Df, Z2 = second-order terms from geo_normal_adjoint or None.  Scatters the grid gradient."""
        net = self.params.nets["geo"]
        nl = len(net) - 2
        Ls, Lf = net[-2], net[-1]
        c = self.cskip
        last = A[nl]
        self.sync_h(dO, self.Df, rows)
        pp = [self.mat("geo_dz0", rows, self.Df, "a", grad=True), self.mat("geo_dz1", rows, self.Df, "a", grad=True)]
        if self.fused_calls:
            # the MLP part of the sweep as ONE C-ABI call (ndjir_geo_backward, csrc/fused_path.cu); the scatter of the
            # grid-feature gradient (with its multi-GPU exchange) stays below
            dgrid = self.mat("geo_dgrid", rows, max(self.Dg, 1), "f") if self.Dg else None
            st = h16.GeoStore()
            for l in range(nl + 1):
                st.acts[l] = A[l].hmat(0, track=False)
            z2 = (h16.HMat * nl)(*[z.hmat(0, track=False) for z in Z2]) if Z2 else None
            dz = (h16.HMat * 2)(pp[0].hmat(0), pp[1].hmat(0))
            self.n_launches += 2 + (3 if dsdf is not None else 0) + 2 * nl - 1 + (2 if self.Dg else 0) - 1
            self.call("ndjir_geo_backward", self.geo_net_desc(), self._grads(net[:nl]), self._grads([Ls]),
                      self._grads([Lf]), rows, st, dO.hmat(0, track=False), (P_(dsdf) if dsdf is not None else None), z2,
                      dz, (dgrid.fptr() if dgrid is not None else None), (dgrid.ldf if dgrid is not None else 0))
            for part, width, off in self._grid_parts():
                tmp = self.buf(f"gg_{part}", rows, width)
                self.copy2d(rows, width, P_(tmp), width, dgrid.fptr(off), dgrid.ldf)
                self._grid_scatter("grad_feature", part, rows, x, [tmp])
            return
        # last layer
        self.wgrad_(rows, Lf, last, dO, 0)
        cur = pp[0]
        self.dgrad(rows, Lf, dO, 0, cur, 0, EPI_MUL_S, H=last, U=(Z2[nl - 1] if Z2 else None))
        if dsdf is not None:
            dsdf_m = Mat(f=dsdf)
            self.wgrad_(rows, Ls, last, dsdf_m, 0)
            self.dgrad(rows, Ls, dsdf_m, 0, cur, 0, EPI_MUL_S, H=last, U=cur)
        dgrid = self.mat("geo_dgrid", rows, max(self.Dg, 1), "f") if self.Dg else None
        skip_wrote = False
        for l in range(nl - 1, -1, -1):
            L = net[l]
            Al = A[l]
            self.wgrad_(rows, L, Al, cur, 0)
            if l > 0:
                is_skip = (l == self.skip)
                n_prev = net[l - 1].N
                nxt = pp[1] if cur is pp[0] else pp[0]
                self.dgrad(rows, L, cur, 0, nxt, 0, EPI_MUL_S, ncols=n_prev, alpha=(c if is_skip else 1.0), H=Al,
                           hscale=(1.0 / c if is_skip else 1.0), U=(Z2[l - 1] if Z2 else None))
                if is_skip and self.Dg:
                    self.dgrad(rows, L, cur, 0, dgrid, 0, EPI_BIAS, row0=n_prev + self.npe, ncols=self.Dg, alpha=c)
                    skip_wrote = True
                cur = nxt
            elif self.Dg:
                self.dgrad(rows, L, cur, 0, dgrid, 0, EPI_ACCUM if skip_wrote else EPI_BIAS, row0=self.npe,
                           ncols=self.Dg)
        for part, width, off in self._grid_parts():
            tmp = self.buf(f"gg_{part}", rows, width)
            self.copy2d(rows, width, P_(tmp), width, dgrid.fptr(off), dgrid.ldf)
            self._grid_scatter("grad_feature", part, rows, x, [tmp])

    # ------------------------------------------------------------------------------------------------
    # generic softplus MLP (heads): forward keeps layer inputs, backward accumulates weight gradients
    # ------------------------------------------------------------------------------------------------
    def mlp_desc(self, name, n_out):
        """POD description of a head for the C-ABI fused path (ndjir_mlp_desc): hidden layers, then the reference's last
        layer as n_out column blocks; planes of W^T (forward) and of W (input gradients) for the tensor-core products"""
        ps, net = self.params, self.params.nets[name]
        nh = len(net) - n_out
        d = h16.MlpDesc()
        d.n_hidden, d.n_out, d.precise = nh, n_out, int(self.precise_fwd)

        def layer(L):
            m = h16.MlpLayer()
            m.K, m.N, m.W, m.ldw, m.bias = L.K, L.N, ps.W(L), L.ldw, ps.b(L)
            m.Wt = ps.WT16(L) if L.N > 8 else h16.NULL_H
            m.Wp = ps.W16(L)
            return m

        for l in range(nh):
            d.hidden[l] = layer(net[l])
        for i, L in enumerate(net[nh:]):
            d.out[i] = layer(L)
        return d

    def _dmat(self, M, col=0, out=False):
        """ndjir_mlp_dmat of a matrix: its fp32 rows when it has them (results) / no planes (operands), else its planes"""
        m = h16.MlpDmat()
        if M.f is not None and (out or M.h is None):
            m.d32, m.ld = M.fptr(col), M.ldf
        else:
            m.dh = M.hmat(col, track=out)
        return m

    def _grads(self, layers):
        ps = self.params
        return (h16.MlpGrad * max(len(layers), 1))(*[h16.MlpGrad(ps.gW(L), ps.gb(L)) for L in layers])

    def mlp_forward(self, name, tag, X, rows, outs):
        """X: input matrix; outs: list of (matrix, column) for the (possibly split) last reference layer."""
        net = self.params.nets[name]
        nh = len(net) - len(outs)
        if self.fused_calls:
            # the whole head as ONE C-ABI call (ndjir_mlp_forward, csrc/fused_path.cu): same products, same buffers
            acts = [self.mat(f"{tag}_h{l}", rows, net[l].N, "a") for l in range(nh)]
            d = self.mlp_desc(name, len(outs))
            out32 = (ctypes.c_void_p * 4)()
            ld_out = (ctypes.c_longlong * 4)()
            outh = (h16.HMat * 4)()
            for i, (Y, ycol) in enumerate(outs):
                if Y.f is not None:
                    out32[i], ld_out[i] = Y.fptr(ycol), Y.ldf
                else:
                    out32[i], ld_out[i], outh[i] = None, 0, Y.hmat(ycol)
            acts_h = (h16.HMat * max(nh, 1))(*[a.hmat(0) for a in acts])
            self.n_launches += nh + len(outs) - 1
            self.call("ndjir_mlp_forward", d, rows, X.hmat(0, track=False), acts_h, out32, ld_out, outh)
            return acts
        acts, cur = [], X
        for l in range(nh):
            L = net[l]
            nxt = self.mat(f"{tag}_h{l}", rows, L.N, "a")
            self.fwd(rows, L, cur, nxt, 0, EPI_SOFTPLUS)
            acts.append(nxt)
            cur = nxt
        for (Y, ycol), L in zip(outs, net[nh:]):
            self.fwd(rows, L, cur, Y, ycol, EPI_BIAS)
        return acts

    def mlp_backward(self, name, tag, X, rows, acts, douts, dX=None, accum_dx=False, dx_cols=None):
        """douts: list of (matrix, column) matching the last (split) layers.  dX[:, :dx_cols] (+)= input gradient."""
        net = self.params.nets[name]
        nh = len(net) - len(douts)
        wmax = max(L.N for L in net[:nh]) if nh else 8
        pp = [self.mat("mlp_dz0", rows, wmax, "a", grad=True), self.mat("mlp_dz1", rows, wmax, "a", grad=True)]
        if self.fused_calls:
            # the whole reverse sweep as ONE C-ABI call (ndjir_mlp_backward, csrc/fused_path.cu)
            d = self.mlp_desc(name, len(douts))
            dys = (h16.MlpDmat * 4)(*[self._dmat(dY, dcol) for dY, dcol in douts])
            acts_h = (h16.HMat * max(nh, 1))(*[a.hmat(0, track=False) for a in acts[:nh]])
            dz = (h16.HMat * 2)(pp[0].hmat(0), pp[1].hmat(0))
            dxm = self._dmat(dX, 0, out=True) if dX is not None else None
            n = 0                      # launches the call makes (products + column sums), for gpu_launches
            for (dY, _), L in zip(douts, net[nh:]):
                n += 1 + (dY.h is None) + (1 if (nh or dX is not None) else 0)
            n += 2 * nh - (1 if (nh and dX is None) else 0)
            if dX is not None and dX.f is not None and 256 < (dx_cols or net[0].K) <= 264:
                n += 1
            self.n_launches += n - 1
            self.call("ndjir_mlp_backward", d, self._grads(net[:nh]), self._grads(net[nh:]), rows, X.hmat(0, track=False),
                      acts_h, dys, dz, dxm, dx_cols or 0, int(accum_dx))
            return
        lastA = acts[-1] if nh else X
        cur = pp[0]
        first = True
        for (dY, dcol), L in zip(douts, net[nh:]):
            self.wgrad_(rows, L, lastA, dY, dcol)
            if nh:
                self.dgrad(rows, L, dY, dcol, cur, 0, EPI_MUL_S, H=lastA, U=(None if first else cur))
            elif dX is not None:
                self.dgrad(rows, L, dY, dcol, dX, 0, EPI_ACCUM if (accum_dx or not first) else EPI_BIAS,
                           ncols=dx_cols or L.K)
            first = False
        for l in range(nh - 1, -1, -1):
            L = net[l]
            Ain = acts[l - 1] if l > 0 else X
            self.wgrad_(rows, L, Ain, cur, 0)
            if l > 0:
                nxt = pp[1] if cur is pp[0] else pp[0]
                self.dgrad(rows, L, cur, 0, nxt, 0, EPI_MUL_S, H=Ain)
                cur = nxt
            elif dX is not None:
                self.dgrad(rows, L, cur, 0, dX, 0, EPI_ACCUM if accum_dx else EPI_BIAS, ncols=dx_cols or L.K)

    # ------------------------------------------------------------------------------------------------
    # sample_points (python/sampler.py:256-299)
    # ------------------------------------------------------------------------------------------------
    def geo_net_desc(self):
        """POD description of the geometric network for the C-ABI fused path (ndjir_geo_net): hidden layers with the
        planes of W^T, the sdf column, skip layer, encoding and grid tables."""
        ps, g = self.params, self.conf.geometric_network
        net = ps.nets["geo"]
        d = h16.GeoNet()
        d.n_hidden = len(net) - 2

        def layer(L, planes):
            m = h16.MlpLayer()
            m.K, m.N, m.W, m.ldw, m.bias = L.K, L.N, ps.W(L), L.ldw, ps.b(L)
            m.Wt = ps.WT16(L) if planes else h16.NULL_H
            m.Wp = ps.W16(L) if planes else h16.NULL_H
            return m

        for l in range(d.n_hidden):
            d.hidden[l] = layer(net[l], True)
        d.sdf = layer(net[-2], False)
        d.feat = layer(net[-1], True)
        d.skip_layer, d.skip_scale, d.pe_bands = self.skip, self.cskip, g.pe_bands
        v = g.voxel
        d.grid_kind = {"voxel": 1, "triplaneline": 2}.get(v.type, 0)
        d.grid_size, d.grid_channels = (v.grid_size, v.feature_size) if d.grid_kind else (0, 0)
        d.grid0 = ps.grid["voxel"].data_ptr() if d.grid_kind == 1 else (ps.grid["triplane"].data_ptr() if d.grid_kind == 2 else None)
        d.grid1 = ps.grid["triline"].data_ptr() if d.grid_kind == 2 else None
        d.precise = int(self.precise_fwd)
        d.use_ste = int(bool(v.use_ste))
        return d

    def _sample_points_c(self, camloc, raydir, stratified_sample, background_sample, mask_sum):
        """sample_points as ONE C-ABI call (ndjir_sample_points_fwd: the round loop runs in the library, csrc/fused_path.cu)
        on the same buffers and scale slots the Python sequencing below uses."""
        r = self.conf.renderer
        B, R, _ = raydir.shape
        NR = B * R
        N0, M, U, Nb = r.n_samples0, r.n_samples1, r.n_upsamples, r.n_bg_samples
        N, Mx = N0 + U * M, max(N0, M)
        cfg = h16.SamplerConfig(N0, M, U, Nb, float(r.sampling_sigmoid_gain),
                                {"intersect_with_aabb": 0, "intersect_with_r_sphere": 1}[r.t_near_far_method], self.rad)
        ws = h16.SamplerWorkspace()
        ws.t_near, ws.t_far, ws.n_hits = (self.buf(k, NR, 1).data_ptr() for k in ("t_near", "t_far", "n_hits"))
        mask = self.buf("mask", NR, 1)
        cur = self.buf("t_a", NR, N + 1)
        ws.sdf_cur = self.buf("smp_sdf_m", NR, N).data_ptr()
        ws.t_pend = self.buf("smp_t_pend", NR, Mx).data_ptr()
        ws.t_new[0], ws.t_new[1] = (self.buf(f"smp_tnew{i}", NR, M).data_ptr() for i in (0, 1))
        self._reserve = NR * Mx
        ws.x = self.buf("smp_x", NR * Mx, 3).data_ptr()
        ws.sdf_pend = self.buf("smp_sdf", NR * Mx, 1).data_ptr()
        A0 = self.mat("smp_A0", NR * Mx, self.din, "fa")
        widest = max(L.K for L in self.params.nets["geo"][1:])
        act = [self.mat(f"geo_pp{i}", NR * Mx, widest, "a") for i in (0, 1)]
        self._reserve = 0
        ws.geo.enc, ws.geo.ld_enc = A0.f.data_ptr(), self.ld0
        gw = sum(w for _, w, _ in self._grid_parts())
        ws.geo.grid_tmp = self.buf("gq_fused", NR * Mx, max(gw, 1)).data_ptr() if gw else None
        ws.geo.ench = A0.hmat(0)
        ws.geo.act[0], ws.geo.act[1] = act[0].hmat(0), act[1].hmat(0)
        x_fg = self.buf("x_fg", NR * N, 3)
        t_bg = self.buf("t_bg", NR, Nb + 1)
        x_bg = self.buf("x_bg", NR * Nb, 4)
        # kernel launches inside the call: bounds, mask, stratified, per round (points, encoding, grid, padding, pack,
        # layers, sdf, placement), final merge, t_far column, points, background
        self.n_launches += 6 + U * (len(self.params.nets["geo"]) + 6 + 2 * len(self._grid_parts()))
        self.call("ndjir_sample_points_fwd", cfg, self.geo_net_desc(), B, R, P_(camloc), P_(raydir), P_(stratified_sample),
                  P_(background_sample), ws, P_(x_fg), P_(cur), P_(x_bg), P_(t_bg), P_(mask),
                  P_(mask_sum) if mask_sum is not None else None)
        self.debug["sampler"] = []
        return (x_fg[:NR * N].view(B, R, N, 3), cur[:NR].reshape(B, R, N + 1, 1), x_bg[:NR * Nb].view(B, R, Nb, 4),
                t_bg[:NR].view(B, R, Nb + 1, 1), mask[:NR].reshape(B, R, 1, 1))

    def sample_points(self, camloc, raydir, stratified_sample, background_sample, mask_sum=None, debug=False):
        """camloc (B,3), raydir (B,R,3), stratified_sample (B,R,N0,1), background_sample (B,R,Nb+1,1) device fp32.
        Returns x_fg (B,R,N,3), t_fg (B,R,N+1,1), x_bg (B,R,Nb,4), t_bg (B,R,Nb+1,1), mask (B,R,1,1)."""
        r = self.conf.renderer
        B, R, _ = raydir.shape
        NR = B * R
        if not getattr(self, "_weights_synced", False):
            self.refresh_transposes()   # standalone call: the registered lo copies of the weights must be current
        if self.fused_calls and not debug:
            return self._sample_points_c(camloc, raydir, stratified_sample, background_sample, mask_sum)
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
        a new graph.  With world_size > 1 the NCCL exchanges (mask sum, sparse grid-gradient all-gather, gradient
        all-reduce) are captured inside the graph: every rank captures and replays the same sequence."""
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
                   backward=True, keep=False, inference=False, obj_mask=None):
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
        if self.h16 and not self._calibrated:
            # First use of the split-fp16 engine: the per-tensor scales are derived from the running maxima of the
            # tensors' PREVIOUS use, so the step is run twice beforehand (results discarded) to seed them.
            self._calibrated = True
            for _ in range(2):
                self.train_step(camloc, raydir, color_gt, rnd, cos_anneal_ratio=cos_anneal_ratio, samples=samples,
                                zero_grad=zero_grad, backward=(backward and zero_grad), inference=inference,
                                obj_mask=obj_mask)
        if zero_grad:
            self.n_launches += 1 + len(ps.grid_grad)
            ps.zero_grad(self.stream())
        self.refresh_transposes()       # the normal pass (forward half) already needs W^T
        if self.h16:
            self.n_launches += 1
            self.scales.update(self.stream())    # last step's maxima -> this step's power-of-two scales
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
        O = self.mat("O", P, Df + 6, "fa")       # [feature | x | normal]: network output, head input, render operand
        A, sdf = self.geo_forward(x_fg, P, "main", store=True, O=O)
        nrm = self.buf("nrm", P, 3)
        GZ, Gin = self.geo_normal(x_fg, P, "main", A, nrm)
        self.copy2d(P, 3, O.fptr(Df), LDO, P_(x_fg), 3)
        self.copy2d(P, 3, O.fptr(Df + 3), LDO, P_(nrm), 3)
        self.sync_h(O, Df + 6, P)
        # ---------------- NeuS alpha, background, compositing ----------------
        alpha_fg = self.buf("alpha_fg", P, 1)
        gain_p = ps.data.data_ptr() + 4 * ps.gain_off
        bgc = conf.background_network
        rows_bg = NR * Nb
        nbg0 = pe_dim(4, bgc.pe_bands0)
        Xbg0 = self.mat("Xbg0", rows_bg, nbg0, "fa")
        self.call("ndjir_positional_encoding", rows_bg, 4, bgc.pe_bands0, P_(x_bg), 4, 1, Xbg0.fptr(), Xbg0.ldf)
        self.sync_h(Xbg0, nbg0, rows_bg)
        Dfb = bgc.feature_size0
        nvpe = pe_dim(3, bgc.pe_bands1)
        Xbg1 = self.mat("Xbg1", rows_bg, Dfb + 4 + 3 + nvpe, "a")
        dens = self.buf("bg_dens", rows_bg, 1)
        acts_bg0 = self.mlp_forward("bg0", "bg0", Xbg0, rows_bg, [(Mat(f=dens), 0), (Xbg1, 0)])
        self.fill_cols(Xbg1, Dfb, P_(x_bg), 4, 4, rows_bg)
        view = self.buf("view", NR, 3)
        self.copy2d(NR, 3, P_(view), 3, P_(raydir), 3, alpha=-1.0)
        self.fill_cols(Xbg1, Dfb + 4, P_(view), 3, 3, rows_bg, rep=Nb)
        vpe_bg = self.buf("vpe_bg", NR, r4(nvpe))
        self.call("ndjir_positional_encoding", NR, 3, bgc.pe_bands1, P_(view), 3, 1, P_(vpe_bg), vpe_bg.shape[1])
        self.fill_cols(Xbg1, Dfb + 7, P_(vpe_bg), vpe_bg.shape[1], nvpe, rows_bg, rep=Nb)
        bgraw = self.buf("bg_raw", rows_bg, 4)
        acts_bg1 = self.mlp_forward("bg1", "bg1", Xbg1, rows_bg, [(Mat(f=bgraw), 0)])
        # one fused kernel per ray: NeuS alpha, background alpha, transmittance scan, weights, VR([feature | x | normal]),
        # background colour (renderer.py:54-91; csrc/render_segment.cu)
        alpha_bg = self.buf("alpha_bg", rows_bg, 1)
        w = self.buf("w", NR, S)
        T = self.buf("T", NR, S)
        colbg = self.buf("colbg", NR, 3)
        pix = self.buf("pix", NR, LDO)
        self.call("ndjir_render_segment_forward", NR, N, Nb, Df + 6, P_(sdf), P_(nrm), 3, P_(raydir), P_(t_fg), gain_p,
                  float(cos_anneal_ratio), P_(maskv), P_(dens), 1, P_(t_bg), P_(bgraw), 4, O.fptr(), LDO, P_(alpha_fg),
                  P_(alpha_bg), P_(w), P_(T), P_(pix), LDO, P_(colbg))
        nhat = self.buf("nhat", NR, 3)
        self.call("ndjir_pixel_normal_forward", NR, P_(pix, Df + 3), LDO, float(r.eps_normal), P_(nhat))
        # ---------------- per-sample heads ----------------
        RAW = self.buf("RAW", P, 16, zero=inference)     # inference leaves column 13 (perturbed base colour) unset
        RAWm = Mat(f=RAW)
        acts = {}
        acts["bc"] = self.mlp_forward("bc", "bc", O, P, [(RAWm, 0)])
        if self.use_ii:
            acts["ii"] = self.mlp_forward("ii", "ii", O, P, [(RAWm, 3)])
        acts["ro"] = self.mlp_forward("ro", "ro", O, P, [(RAWm, 4)])
        acts["sp"] = self.mlp_forward("sp", "sp", O, P, [(RAWm, 6)])
        Xpl = None
        if self.use_pl:
            plc = conf.photogrammetric_light_network
            npl = pe_dim(3, plc.pe_bands)
            use_inv = bool(plc.use_inverse_distance)
            Xpl = self.mat("Xpl", P, Df + 6 + npl + int(use_inv), "a", slot="O")      # [O | PE(view) | 1/d^2], same scale as O
            self.copy_cols(Xpl, O, Df + 6, P)
            vpe_pl = self.buf("vpe_pl", NR, r4(npl))
            self.call("ndjir_positional_encoding", NR, 3, plc.pe_bands, P_(view), 3, 1, P_(vpe_pl), vpe_pl.shape[1])
            self.fill_cols(Xpl, Df + 6, P_(vpe_pl), vpe_pl.shape[1], npl, P, rep=N)
            if use_inv:
                invd = self.buf("inv_sq_dist", P, 1)
                self.call("ndjir_inv_sq_dist", P, R * N, P_(x_fg), P_(camloc), P_(invd), 1)
                self.fill_cols(Xpl, Df + 6 + npl, P_(invd), 1, 1, P)
            acts["pl"] = self.mlp_forward("pl", "pl", Xpl, P, [(RAWm, 12)])
        # ---------------- perturbed colour branch (renderer.py:187-193) ----------------
        # It only feeds the base-colour prior of the loss: image rendering (inference=True, forward only) skips the
        # second geometric-network evaluation and reads a zero base colour there.
        assert not (inference and backward), "inference=True is forward only"
        if not inference:
            G = conf.geometric_network.voxel.grid_size
            x_ptb = self.buf("x_ptb", P, 3)
            self.copy2d(P, 3, P_(x_ptb), 3, P_(x_fg), 3)
            self.copy2d(P, 3, P_(x_ptb), 3, P_(rnd["perturb"]), 3, alpha=math.sqrt(3) * 2 * self.rad / G, accum=1)
            Op = self.mat("O_ptb", P, Df + 6, "fa")
            Ap, _ = self.geo_forward(x_ptb, P, "ptb", store=True, O=Op)
            self.copy2d(P, 3, Op.fptr(Df), LDO, P_(x_ptb), 3)
            self.sync_h(Op, Df + 3, P)
            acts["bcp"] = self.mlp_forward("bc", "bcp", Op, P, [(RAWm, 13)])
        # ---------------- material attributes + per-sample losses ----------------
        ro_c, sp_c = conf.roughness_network, conf.specular_reflectance_network
        # filament + remap: specular reflectance = 0.16 sigmoid(h)^2 with the literal 0.16 of network.py:503-504
        # (upper_bound_scale only enters the other branch, :506, which __init__ rejects)
        cfg10 = [ro_c.lower_bound, ro_c.prior_value, sp_c.prior_value, 0.16, ps.pl_gain,
                 tr.eikonal_weight, tr.base_color_prior_weight, tr.roughness_prior_weight,
                 tr.specular_reflectance_prior_weight,
                 float(int(bool(tr.base_color_prior_sym_backward)) + (0 if conf.diffuse_brdf.entangle else 2)
                       + (0 if self.use_ii else 4) + (0 if self.use_pl else 8))]
        ATT = self.buf("ATT", P, 12)
        self.call("ndjir_sample_attributes_forward", P, N, P_(RAW), P_(ATT), P_(nrm), 3, P_(maskv), cfg10, P_(losses))
        # mask loss (loss.py:108-116): obj_mask_pred = sum_i alpha_i T_i (renderer.py:183-185) is the volume-rendering
        # reduction of a constant-1 attribute: spare column 9 of ATT carries it through the forward reduction and, with
        # d_attpix[:, 9] = d loss / d obj_mask_pred, through the weight gradients of the compositing backward
        mask_term = tr.mask_weight > 0.0
        if mask_term:
            obj_mask = rnd.get("obj_mask") if obj_mask is None else obj_mask
            if obj_mask is None:
                raise ValueError("train.mask_weight > 0 needs obj_mask (B, R, 1)")
            self.call("ndjir_copy2d", P, 1, P_(ATT, 9), 12, P_(self.one), 0, 1, 1.0, 0)
        attpix = self.buf("attpix", NR, 12)
        self.call("ndjir_volume_render_forward", NR, N, 12, P_(w), S, P_(ATT), 12, P_(attpix), 12)
        # ---------------- light directions, environment light, soft visibility ----------------
        dirs_u = self.buf("dirs_u", NR * M, 3)
        dirs_s = self.buf("dirs_s", NR * M, 3)
        self.call("ndjir_sample_uniform_directions", NR * M, P_(dirs_u), P_(nhat), P_(rnd["diffuse_cdf_the"]),
                  P_(rnd["diffuse_cdf_phi"]), NR, M, nt, 2 * nt, 0.0)
        rho = self.buf("rho_pix", NR, 1)
        self.copy2d(NR, 1, P_(rho), 1, P_(attpix, 1), 12)
        if conf.specular_brdf.sampling == "importance":
            self.call("ndjir_sample_importance_directions", NR * M, P_(dirs_s), P_(nhat), P_(rnd["specular_cdf_the"]),
                      P_(rnd["specular_cdf_phi"]), P_(rho), NR, M, nt, 2 * nt, 0.0)
        else:       # renderer.py:137-140 (config/uniform_sampling_on_sepcular.yaml)
            self.call("ndjir_sample_uniform_directions", NR * M, P_(dirs_s), P_(nhat), P_(rnd["specular_cdf_the"]),
                      P_(rnd["specular_cdf_phi"]), NR, M, nt, 2 * nt, 0.0)
        elc, svc = conf.environment_light_network, conf.soft_visibility_light_network
        rows_d = NR * 2 * M
        nel = pe_dim(3, elc.pe_bands)
        Xel = self.mat("Xel", rows_d, nel, "fa")
        ldel = Xel.ldf
        # rows (set, r, j): the diffuse set first, then the specular set
        for s_, dirs in enumerate((dirs_u, dirs_s)):
            self.call("ndjir_positional_encoding", NR * M, 3, elc.pe_bands, P_(dirs), 3, 1, Xel.fptr(0, s_ * NR * M), ldel)
        self.sync_h(Xel, nel, rows_d)
        elraw = self.buf("el_raw", rows_d, 4)
        acts["el"] = self.mlp_forward("el", "el", Xel, rows_d, [(Mat(f=elraw), 0)])
        nsv = pe_dim(3, svc.pe_bands)
        assert nsv == nel
        Xsv = self.mat("Xsv", rows_d, Df + 6 + nsv, "a")
        for s_ in range(2):
            self.fill_cols(Xsv, 0, P_(pix), LDO, Df + 3, NR * M, rep=M, drow=s_ * NR * M)            # f_pix | x_pix
            self.fill_cols(Xsv, Df + 3, P_(nhat), 3, 3, NR * M, rep=M, drow=s_ * NR * M)           # nhat
        self.fill_cols(Xsv, Df + 6, Xel.fptr(), ldel, nsv, rows_d)                                 # PE(omega)
        svraw = self.buf("sv_raw", rows_d, 4)
        acts["sv"] = self.mlp_forward("sv", "sv", Xsv, rows_d, [(Mat(f=svraw), 0)])
        # ---------------- shading + colour loss ----------------
        cfg5 = [r.eps_dot, conf.specular_brdf.weight, inv_rays,
                float(int(bool(conf.diffuse_brdf.entangle)) + (2 if conf.specular_brdf.sampling == "uniform" else 0)
                      + (0 if self.use_pl else 4)),
                0.0 if tr.rgb_loss == "l1" else 1.0]
        color = self.buf("color", NR, 3)
        ray_w = None
        if mask_term:     # loss.py:63-65: the colour loss over the object's rays only, / (sum(obj_mask) + 1e-5)
            self.call("ndjir_colsum", NR, 1, P_(scal, 2), P_(obj_mask), 1, 1.0)
            if self.world_size > 1:
                allreduce_mask_sum(scal[0, 2:3], self.pg)
            ray_w = self.buf("ray_loss_w", NR, 1)
            self.call("ndjir_ray_loss_weights", NR, P_(obj_mask), P_(scal, 2), float(NR * self.world_size), P_(ray_w))
        self.call("ndjir_shade_forward_weighted", NR, M, P_(nhat), P_(attpix), P_(raydir), P_(dirs_u), P_(dirs_s),
                  P_(elraw), 4, P_(svraw), 4, P_(colbg), P_(color_gt), cfg5, P_(ray_w) if ray_w is not None else None,
                  P_(color), P_(losses))
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
        if mask_term:
            self.call("ndjir_mask_loss", NR, N, P_(attpix), 12, 9, P_(alpha_fg), P_(maskv), P_(obj_mask), mask_sum,
                      float(tr.mask_weight), P_(losses), None, 0, None)
        if self.world_size > 1:
            allreduce_mask_sum(losses[0, :N_LOSSES], self.pg)     # report the loss of the union of all ranks' rays
        if keep:
            self.debug.update(dict(O=O.f, sdf=sdf, nrm=nrm, alpha_fg=alpha_fg, alpha_bg=alpha_bg, w=w, T=T, pix=pix,
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
        self.call("ndjir_shade_backward_weighted", NR, M, P_(nhat), P_(attpix), P_(raydir), P_(dirs_u), P_(dirs_s),
                  P_(elraw), 4, P_(svraw), 4, P_(colbg), P_(color_gt), cfg5, P_(ray_w) if ray_w is not None else None,
                  P_(d_el), P_(d_sv), P_(d_attpix), P_(d_nhat), P_(d_colbg))
        if mask_term:
            dalpha_missed = self.buf("dalpha_missed", P, 1)
            self.call("ndjir_mask_loss", NR, N, P_(attpix), 12, 9, P_(alpha_fg), P_(maskv), P_(obj_mask), mask_sum,
                      float(tr.mask_weight), None, P_(d_attpix), 12, P_(dalpha_missed))
        # environment light: parameters only (directions carry no gradient, sampler.py:391)
        self.mlp_backward("el", "el", Xel, rows_d, acts["el"], [(Mat(f=d_el), 0)])
        # soft visibility: input gradient -> per-ray sums
        dXsv = self.mat("dXsv", rows_d, Df + 6, "f")
        self.mlp_backward("sv", "sv", Xsv, rows_d, acts["sv"], [(Mat(f=d_sv), 0)], dX=dXsv, dx_cols=Df + 6)
        dpix = self.buf("dpix", NR, LDO, zero=True)
        self.call("ndjir_group_sum", NR, M, Df + 6, P_(dpix), LDO, dXsv.fptr(), dXsv.ldf, 0)
        self.call("ndjir_group_sum", NR, M, Df + 6, P_(dpix), LDO, dXsv.fptr(0, NR * M), dXsv.ldf, 1)
        # dpix columns Df+3:Df+6 currently hold d nhat from the visibility input; add the shading part, then
        # turn d nhat into d n_pix
        self.copy2d(NR, 3, P_(d_nhat), 3, P_(dpix, Df + 3), LDO, accum=1)
        self.call("ndjir_pixel_normal_backward", NR, P_(pix, Df + 3), LDO, float(r.eps_normal), P_(d_nhat),
                  P_(dpix, Df + 3), LDO, 0)
        # one fused kernel per ray: weight gradients of both reductions and the background colour, the suffix scan of the
        # compositing backward, NeuS alpha backward (dsdf, dnormal, dgain) and background density backward
        dO = self.mat("dO", P, Df + 6, "fa", grad=True)
        dATT = self.buf("dATT", P, 12)
        d_bgraw = self.buf("d_bgraw", rows_bg, 4)
        d_dens = self.buf("d_dens", rows_bg, 1)
        dsdf = self.buf("dsdf", P, 1)
        g_gain = ps.grad.data_ptr() + 4 * ps.gain_off
        dw = self.buf("dw", NR, S) if keep else None
        dalpha_fg = self.buf("dalpha_fg", P, 1) if keep else None
        dalpha_bg = self.buf("dalpha_bg", rows_bg, 1) if keep else None
        self.call("ndjir_render_segment_backward", NR, N, Nb, Df + 6, 12, P_(sdf), P_(nrm), 3, P_(raydir), P_(t_fg),
                  gain_p, float(cos_anneal_ratio), P_(maskv), P_(dens), 1, P_(t_bg), P_(bgraw), 4, O.fptr(), LDO,
                  P_(ATT), 12, P_(w), P_(T), P_(dpix), LDO, P_(d_attpix), 12, P_(d_colbg), dO.fptr(), LDO, P_(dATT), 12,
                  P_(d_bgraw), 4, P_(d_dens), 1, P_(dsdf), dO.fptr(Df + 3), LDO, g_gain,
                  P_(dw) if keep else None, P_(dalpha_fg) if keep else None, P_(dalpha_bg) if keep else None)
        if mask_term:   # rays that miss the bounds: obj_mask_pred = sum_i alpha_fg_i reaches the SDF network through alpha alone
            self.call("ndjir_neus_alpha_backward", P, N, P_(dalpha_missed), P_(sdf), P_(nrm), 3, P_(raydir), P_(t_fg),
                      gain_p, float(cos_anneal_ratio), P_(dsdf), dO.fptr(Df + 3), LDO, g_gain)
        # material heads
        dRAW = self.buf("dRAW", P, 16)
        dRAWm = Mat(f=dRAW)
        self.call("ndjir_sample_attributes_backward", P, N, P_(RAW), P_(dATT), P_(nrm), 3, P_(maskv), cfg10, inv_denorm,
                  P_(dRAW), dO.fptr(Df + 3), LDO)
        self.mlp_backward("bc", "bc", O, P, acts["bc"], [(dRAWm, 0)], dX=dO, accum_dx=True, dx_cols=Df + 3)
        for name, col in (("ii", 3), ("ro", 4), ("sp", 6)):
            if name in acts:
                self.mlp_backward(name, name, O, P, acts[name], [(dRAWm, col)], dX=dO, accum_dx=True, dx_cols=Df + 6)
        if self.use_pl:
            self.mlp_backward("pl", "pl", Xpl, P, acts["pl"], [(dRAWm, 12)], dX=dO, accum_dx=True, dx_cols=Df + 6)
        # background networks
        dXbg1 = self.mat("dXbg1", rows_bg, Dfb, "a", grad=True)
        self.mlp_backward("bg1", "bg1", Xbg1, rows_bg, acts_bg1, [(Mat(f=d_bgraw), 0)], dX=dXbg1, dx_cols=Dfb)
        self.mlp_backward("bg0", "bg0", Xbg0, rows_bg, acts_bg0, [(Mat(f=d_dens), 0), (dXbg1, 0)])
        # geometric network: second-order terms from d L / d normal, then the standard sweep
        nbar = self.buf("nbar", P, 3)
        self.copy2d(P, 3, P_(nbar), 3, dO.fptr(Df + 3), LDO)
        Z2 = self.geo_normal_adjoint(x_fg, P, "main", A, GZ, Gin, nbar)
        self.geo_backward(x_fg, P, "main", A, dsdf, dO, Z2)
        # perturbed branch
        dOp = self.mat("dO_ptb", P, Df + 6, "fa", grad=True)
        self.mlp_backward("bc", "bcp", Op, P, acts["bcp"], [(dRAWm, 13)], dX=dOp, dx_cols=Df + 3)
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
            self.debug.update(dict(dO=dO.f, dsdf=dsdf, dw=dw, dRAW=dRAW, d_attpix=d_attpix, dpix=dpix, nbar=nbar,
                                   dalpha_fg=dalpha_fg, dalpha_bg=dalpha_bg, d_el=d_el, d_sv=d_sv))
        self._weights_synced = False
        return losses[0, :N_LOSSES]
