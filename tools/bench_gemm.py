"""Times the MLP product kernel at the default.yaml shapes (P = 262144 sample rows) for both paths."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ndjir_b200 import _lib

P = 262144
EPI_BIAS, EPI_SOFTPLUS, EPI_ACCUM, EPI_MUL_S, EPI_ADJ, EPI_ATOMIC = range(6)
def t(*shape): return torch.randn(*shape, device="cuda") * 0.05
cases = []
A, W, C, H, U, C2, b = t(P, 256), t(256, 256), t(P, 256), t(P, 256).abs(), t(P, 256), t(P, 256), t(256)
gW = torch.zeros(256, 256, device="cuda")
one = 0
def call(M, N, K, A_, ars, acs, B_, brs, bcs, C_, ldc, epi, split=1, Hh=None, Uu=None, C2_=None):
    _lib.call("ndjir_gemm", M, N, K, A_, ars, acs, B_, brs, bcs, C_, ldc, b, 1.0, 1.0, 100.0, Hh, 256, 1.0, Uu, 256, C2_, 256, split, epi, 0)
shapes = {
  "fwd_softplus 256x256": lambda: call(P, 256, 256, A, 256, 1, W, 256, 1, C, 256, EPI_SOFTPLUS),
  "dgrad_mul_s 256x256": lambda: call(P, 256, 256, A, 256, 1, W, 1, 256, C, 256, EPI_MUL_S, Hh=H, Uu=U),
  "adjoint 256x256": lambda: call(P, 256, 256, A, 256, 1, W, 256, 1, C, 256, EPI_ADJ, Hh=H, Uu=U, C2_=C2),
  "wgrad 256x256": lambda: call(256, 256, P, A, 1, 256, C, 256, 1, gW, 256, EPI_ATOMIC, split=74),
}
out = []
import sys as _s
if "--presplit" in _s.argv:
    W_lo = torch.empty_like(W)
    _lib.call("ndjir_split_lo", W.numel(), W_lo, W, 0)
    def call(M, N, K, A_, ars, acs, B_, brs, bcs, C_, ldc, epi, split=1, Hh=None, Uu=None, C2_=None):  # noqa: F811
        _lib.call("ndjir_gemm_presplit", M, N, K, A_, ars, acs, B_, W_lo if B_ is W else None, brs, bcs, C_, ldc, b, 1.0,
                  1.0, 100.0, Hh, 256, 1.0, Uu, 256, C2_, 256, split, epi, 0)
    for pre in (0, 1, 0, 1):
        _lib.call("ndjir_set_option", "mlp_presplit", pre)
        for name, fn in shapes.items():
            for _ in range(3): fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(json.dumps(dict(name=name, presplit=pre, ms=ms, tflops=2.0 * P * 256 * 256 / ms / 1e9)), flush=True)
    _lib.call("ndjir_set_option", "mlp_presplit", 1)
    _s.exit(0)
modes = [(1, d) for d in range(8)] if "--dbg" in _s.argv else ([(1, 0), (2, 0)] if "--pair" in _s.argv else [(1, 0), (0, 0)])
for tc, dbg in modes:
    _lib.call("ndjir_set_option", "mlp_tensor_cores", 1 if tc else 0)
    _lib.call("ndjir_set_option", "mlp_cta_pair", 1 if tc == 2 else 0)
    _lib.call("ndjir_set_option", "mlp_dbg", dbg)
    for name, fn in shapes.items():
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        r = dict(name=name, tc=tc, dbg=dbg, ms=ms, tflops=2.0 * P * 256 * 256 / ms / 1e9)
        out.append(r); print(json.dumps(r), flush=True)
_lib.call("ndjir_set_option", "mlp_tensor_cores", 1)
_lib.call("ndjir_set_option", "mlp_dbg", 0)
_lib.call("ndjir_set_option", "mlp_cta_pair", 0)
