"""Experiment: tcgen05 product kernel at the first-layer shapes (N = 43 outputs / K = 43 inputs) vs regular widths."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndjir_b200 import _lib
P = 262144
def t(*shape): return torch.randn(*shape, device="cuda") * 0.05
b = t(256)
def bench(M, N, K, A, ars, acs, B, brs, bcs, C, ldc, epi):
    fn = lambda: _lib.call("ndjir_gemm", M, N, K, A, ars, acs, B, brs, bcs, C, ldc, b, 1.0, 1.0, 100.0, None, 0, 1.0, None, 0, None, 0, 1, epi, 0)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
A256 = t(P, 256)
for N in (43, 48, 64, 128, 256):
    ld = (N + 3) // 4 * 4
    Wt = t(256, ld)            # B(k, n) = Wt[k, n]: MN-major
    C = t(P, ld)
    print(f"dX = dZ(Px256) W^T -> N={N:3d} (B MN-major, ldb={ld}): {bench(P, N, 256, A256, 256, 1, Wt, ld, 1, C, ld, 0):.3f} ms", flush=True)
    Wk = t(N, 256)             # B(k, n) = Wk[n, k]: K-major
    print(f"                          N={N:3d} (B K-major):              {bench(P, N, 256, A256, 256, 1, Wk, 1, 256, C, ld, 0):.3f} ms", flush=True)
C256 = t(P, 256)
for K in (43, 48, 64, 128, 256):
    ld = (K + 3) // 4 * 4
    A = t(P, ld)
    W = t(K, 256)
    print(f"Z = X(Px{K}) W -> 256 (softplus), lda={ld}: {bench(P, 256, K, A, ld, 1, W, 256, 1, C256, 256, 1):.3f} ms", flush=True)
