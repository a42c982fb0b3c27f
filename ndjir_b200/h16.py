"""Host side of the split-fp16 MLP engine (include/ndjir_b200.h, "split-fp16 MLP engine"): ctypes mirrors of ndjir_hmat /
ndjir_gemm_h_desc, split-tensor buffers and the per-tensor scale slots.  torch only provides the device memory."""
import ctypes

import torch

from . import _lib

EPI_BIAS, EPI_SOFTPLUS, EPI_ACCUM, EPI_MUL_S, EPI_ADJ, EPI_ATOMIC = range(6)


class HMat(ctypes.Structure):
    _fields_ = [("hi", ctypes.c_void_p), ("lo", ctypes.c_void_p), ("ld", ctypes.c_longlong),
                ("scale", ctypes.c_void_p), ("amax", ctypes.c_void_p)]


class GemmHDesc(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int), ("mn_major", ctypes.c_int),
                ("epilogue", ctypes.c_int), ("precise", ctypes.c_int), ("split_k", ctypes.c_int),
                ("alpha", ctypes.c_float), ("out_scale", ctypes.c_float), ("beta", ctypes.c_float),
                ("hscale", ctypes.c_float),
                ("A", HMat), ("B", HMat),
                ("A32", ctypes.c_void_p), ("a_rs", ctypes.c_longlong), ("a_cs", ctypes.c_longlong),
                ("B32", ctypes.c_void_p), ("b_rs", ctypes.c_longlong), ("b_cs", ctypes.c_longlong),
                ("C", ctypes.c_void_p), ("ldc", ctypes.c_longlong), ("Ch", HMat),
                ("C2", ctypes.c_void_p), ("ldc2", ctypes.c_longlong), ("C2h", HMat),
                ("bias", ctypes.c_void_p),
                ("H", ctypes.c_void_p), ("ldh", ctypes.c_longlong), ("Hh", HMat),
                ("U", ctypes.c_void_p), ("ldu", ctypes.c_longlong), ("Uh", HMat),
                ("colsum", ctypes.c_void_p)]


# ---- PODs of the fused-path entry points (include/ndjir_b200.h, "fused path behind the C ABI") ----
MAX_MLP_LAYERS = 16


class MlpLayer(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int), ("N", ctypes.c_int), ("W", ctypes.c_void_p), ("ldw", ctypes.c_longlong),
                ("bias", ctypes.c_void_p), ("Wt", HMat), ("Wp", HMat)]


class MlpDesc(ctypes.Structure):
    _fields_ = [("n_hidden", ctypes.c_int), ("n_out", ctypes.c_int), ("hidden", MlpLayer * MAX_MLP_LAYERS),
                ("out", MlpLayer * 4), ("precise", ctypes.c_int)]


class MlpGrad(ctypes.Structure):
    _fields_ = [("gW", ctypes.c_void_p), ("gb", ctypes.c_void_p)]


class MlpDmat(ctypes.Structure):
    _fields_ = [("d32", ctypes.c_void_p), ("ld", ctypes.c_longlong), ("dh", HMat)]


class GeoNet(ctypes.Structure):
    _fields_ = [("n_hidden", ctypes.c_int), ("hidden", MlpLayer * MAX_MLP_LAYERS), ("sdf", MlpLayer),
                ("skip_layer", ctypes.c_int), ("skip_scale", ctypes.c_float), ("pe_bands", ctypes.c_int),
                ("grid_kind", ctypes.c_int), ("grid_size", ctypes.c_int), ("grid_channels", ctypes.c_int),
                ("grid0", ctypes.c_void_p), ("grid1", ctypes.c_void_p), ("precise", ctypes.c_int), ("feat", MlpLayer),
                ("use_ste", ctypes.c_int)]


class GeoScratch(ctypes.Structure):
    _fields_ = [("enc", ctypes.c_void_p), ("ld_enc", ctypes.c_longlong), ("grid_tmp", ctypes.c_void_p),
                ("ench", HMat), ("act", HMat * 2)]


class GeoStore(ctypes.Structure):
    _fields_ = [("enc", ctypes.c_void_p), ("ld_enc", ctypes.c_longlong), ("grid_tmp", ctypes.c_void_p),
                ("acts", HMat * (MAX_MLP_LAYERS + 1))]


class GeoNormalWs(ctypes.Structure):
    _fields_ = [("gz", HMat * MAX_MLP_LAYERS), ("g_in", ctypes.c_void_p), ("grid_tmp", ctypes.c_void_p),
                ("ones", ctypes.c_void_p)]


class SamplerConfig(ctypes.Structure):
    _fields_ = [("n_samples0", ctypes.c_int), ("n_samples1", ctypes.c_int), ("n_upsamples", ctypes.c_int),
                ("n_bg_samples", ctypes.c_int), ("sampling_sigmoid_gain", ctypes.c_float), ("bounds", ctypes.c_int),
                ("radius", ctypes.c_float)]


class SamplerWorkspace(ctypes.Structure):
    _fields_ = [("t_near", ctypes.c_void_p), ("t_far", ctypes.c_void_p), ("n_hits", ctypes.c_void_p),
                ("sdf_cur", ctypes.c_void_p), ("t_pend", ctypes.c_void_p), ("t_new", ctypes.c_void_p * 2),
                ("x", ctypes.c_void_p), ("sdf_pend", ctypes.c_void_p), ("geo", GeoScratch)]


class Scales:
    """Device arrays scale[n], amax[n] (+ a flag word): one slot per split tensor.  update() turns the running maxima
    of the tensors' last use into the power-of-two scales of their next use (delayed scaling)."""

    def __init__(self, device, n=512, target_log2=10):
        self.scale = torch.ones(n, dtype=torch.float32, device=device)
        self.amax = torch.zeros(n, dtype=torch.float32, device=device)
        self.flags = torch.zeros(4, dtype=torch.int32, device=device)
        self.names = {}
        self.target_log2 = target_log2

    def slot(self, name, init=1.0):
        i = self.names.get(name)
        if i is None:
            i = self.names[name] = len(self.names)
            if i >= self.scale.numel():
                raise _lib.NdjirError("out of scale slots")
            self.scale[i] = init
        return i

    def update(self, stream):
        _lib.call("ndjir_scale_update", len(self.names), self.scale, self.amax, self.flags, self.target_log2, stream)


class HBuf:
    """(rows, ld) split-fp16 matrix: two fp16 planes in one allocation [2, rows, ld]; ld is a multiple of 64 halfs so
    that every row chunk is a whole TMA box row and the weight-gradient products can fetch 3-D boxes."""

    def __init__(self, rows, cols, device, scales=None, name=None, ld=None, init=1.0):
        self.rows, self.cols = rows, cols
        self.ld = ld if ld is not None else (cols + 63) // 64 * 64
        # zeroed once with the library's own fill (the planes viewed as rows * ld floats): padding columns stay zero
        self.t = torch.empty((2, rows, self.ld), dtype=torch.float16, device=device)
        if self.t.is_cuda:
            _lib.call("ndjir_fill", rows * self.ld, self.t.data_ptr(), 0.0, torch.cuda.current_stream().cuda_stream)
        else:
            self.t.zero_()
        self.scales = scales
        self.slot = scales.slot(name, init) if (scales is not None and name is not None) else None

    def hmat(self, col=0, row=0, track=True):
        """ndjir_hmat view starting at (row, col)."""
        off = 2 * (row * self.ld + col)
        hi = self.t.data_ptr() + off
        lo = hi + 2 * self.rows * self.ld
        sc = am = None
        if self.slot is not None:
            sc = self.scales.scale.data_ptr() + 4 * self.slot
            am = (self.scales.amax.data_ptr() + 4 * self.slot) if track else None
        return HMat(hi, lo, self.ld, sc, am)

    def pack(self, src, stream, cols=None, col=0, rep=1, alpha=1.0, rows=None):
        """self[:, col:col+cols] = split(alpha * src[r // rep]) from an fp32 tensor (rows_src, >= cols)."""
        cols = src.shape[1] if cols is None else cols
        rows = self.rows if rows is None else rows
        _lib.call("ndjir_pack_h", rows, cols, src, src.stride(0), rep, alpha, self.hmat(col), stream)

    def unpack(self, stream, cols=None, rows=None):
        cols = self.cols if cols is None else cols
        rows = self.rows if rows is None else rows
        out = torch.empty((rows, cols), dtype=torch.float32, device=self.t.device)
        _lib.call("ndjir_unpack_h", rows, cols, self.hmat(track=False), out, cols, stream)
        return out


NULL_H = HMat(None, None, 0, None, None)


def gemm_h(stream, M, N, K, epi, A=None, B=None, mn_major=False, precise=False, split_k=1, alpha=1.0, out_scale=1.0,
           beta=100.0, hscale=1.0, A32=None, a_rs=0, a_cs=1, B32=None, b_rs=0, b_cs=1, C=None, ldc=0, Ch=None,
           C2=None, ldc2=0, C2h=None, bias=None, H=None, ldh=0, Hh=None, U=None, ldu=0, Uh=None, colsum=None):
    """One product of the split-fp16 engine.  A, B, Ch, C2h, Hh, Uh are HMat views; C, C2, H, U, bias, A32, B32 device
    addresses (int) of fp32 data."""
    d = GemmHDesc()
    d.M, d.N, d.K = M, N, K
    d.mn_major, d.epilogue, d.precise, d.split_k = int(mn_major), epi, int(precise), split_k
    d.alpha, d.out_scale, d.beta, d.hscale = alpha, out_scale, beta, hscale
    d.A = A if A is not None else NULL_H
    d.B = B if B is not None else NULL_H
    d.A32, d.a_rs, d.a_cs = A32, a_rs, a_cs
    d.B32, d.b_rs, d.b_cs = B32, b_rs, b_cs
    d.C, d.ldc = C, ldc
    d.Ch = Ch if Ch is not None else NULL_H
    d.C2, d.ldc2 = C2, ldc2
    d.C2h = C2h if C2h is not None else NULL_H
    d.bias = bias
    d.H, d.ldh = H, ldh
    d.Hh = Hh if Hh is not None else NULL_H
    d.U, d.ldu = U, ldu
    d.Uh = Uh if Uh is not None else NULL_H
    d.colsum = colsum
    _lib.call("ndjir_gemm_h", d, stream)
