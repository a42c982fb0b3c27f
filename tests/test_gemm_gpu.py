"""GPU check of the MLP product engine (ndjir_gemm): the tcgen05 3xTF32 path and the fp32 FFMA path against float64
numpy for every operand layout the MLP passes use (forward: A K-major x B MN-major; dgrad: K-major x K-major; wgrad:
MN-major x MN-major with split-K atomics) and every fused epilogue.  Bar: 1e-5 max-norm relative (the forward
tolerance of BASELINE.json)."""
import numpy as np
import pytest
import torch

from ndjir_b200 import _lib

pytestmark = pytest.mark.gpu
EPI_BIAS, EPI_SOFTPLUS, EPI_ACCUM, EPI_MUL_S, EPI_ADJ, EPI_ATOMIC = range(6)


def r4(n):
    return (n + 3) // 4 * 4


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def softplus(x, beta=100.0):
    z = beta * x
    return (np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))) / beta


def run(M, N, K, layout, epi, tc, split_k=1, seed=0, mask_hi=0, presplit=0, ret_out=False):
    rng = np.random.RandomState(seed)
    A = rng.randn(M, K).astype(np.float32)
    B = (rng.randn(K, N) / np.sqrt(K)).astype(np.float32)
    if epi == EPI_SOFTPLUS:
        B = (B * np.float32(0.05)).astype(np.float32)
    # storage (padding holds NaN: nothing outside the logical extents may be read into a product)
    if layout[0] == "k":      # A row-major (M, lda)
        lda = r4(K) + 4
        As = np.full((M, lda), np.nan, np.float32); As[:, :K] = A
        a_rs, a_cs = lda, 1
    else:                     # A stored transposed (K, lda): contiguous along m
        lda = r4(M) + 4
        As = np.full((K, lda), np.nan, np.float32); As[:, :M] = A.T
        a_rs, a_cs = 1, lda
    if layout[1] == "n":      # B row-major (K, ldb): contiguous along n
        ldb = r4(N) + 4
        Bs = np.full((K, ldb), np.nan, np.float32); Bs[:, :N] = B
        b_rs, b_cs = ldb, 1
    else:                     # B stored (N, ldb): contiguous along k
        ldb = r4(K) + 4
        Bs = np.full((N, ldb), np.nan, np.float32); Bs[:, :K] = B.T
        b_rs, b_cs = 1, ldb
    ldc = r4(N) + 4
    C0 = rng.randn(M, ldc).astype(np.float32)
    bias = rng.randn(N).astype(np.float32) * 0.1
    H = np.abs(rng.randn(M, ldc)).astype(np.float32) * 0.02
    U = rng.randn(M, ldc).astype(np.float32)
    acc = A.astype(np.float64) @ B.astype(np.float64)
    alpha, out_scale, hscale = 0.7, 0.9, 1.3
    s = 1.0 - np.exp(-100.0 * hscale * H[:, :N].astype(np.float64))
    want2 = None
    if epi == EPI_BIAS:
        want = alpha * acc + bias
    elif epi == EPI_SOFTPLUS:
        want = out_scale * softplus(acc + bias)
    elif epi in (EPI_ACCUM, EPI_ATOMIC):
        want = C0[:, :N] + alpha * acc
    elif epi == EPI_MUL_S:
        want = alpha * acc * s + U[:, :N]
    elif epi == EPI_ADJ:
        want = acc * U[:, :N] * 100.0 * (1 - s)
        want2 = out_scale * acc * s
    dA, dB, dC, dH, dU = dev(As), dev(Bs), dev(C0), dev(H), dev(U)
    dC2 = torch.zeros_like(dC)
    dbias = dev(bias)
    _lib.call("ndjir_set_option", "mlp_tensor_cores", int(tc))
    _lib.call("ndjir_set_option", "mlp_mask_hi", int(mask_hi))
    dB_lo = None
    if presplit:      # the weight operand comes with its pre-split lo copy (two TMA tiles, no B transform)
        dB_lo = torch.empty_like(dB)
        _lib.call("ndjir_split_lo", dB.numel(), dB_lo, dB, 0)
    try:
        if presplit:
            _lib.call("ndjir_gemm_presplit", M, N, K, dA, a_rs, a_cs, dB, dB_lo, b_rs, b_cs, dC, ldc, dbias, alpha,
                      out_scale, 100.0, dH, ldc, hscale, dU, ldc, dC2, ldc, split_k, epi, 0)
        else:
            _lib.call("ndjir_gemm", M, N, K, dA, a_rs, a_cs, dB, b_rs, b_cs, dC, ldc, dbias, alpha, out_scale, 100.0,
                      dH, ldc, hscale, dU, ldc, dC2, ldc, split_k, epi, 0)
        torch.cuda.synchronize()
    finally:
        _lib.call("ndjir_set_option", "mlp_tensor_cores", 1)
        _lib.call("ndjir_set_option", "mlp_mask_hi", 0)
    got = dC.cpu().numpy().astype(np.float64)
    err = np.abs(got[:, :N] - want).max() / np.abs(want).max()
    # columns beyond N must be untouched
    assert np.array_equal(got[:, N:], C0[:, N:].astype(np.float64)), "wrote outside the tile"
    if want2 is not None:
        err = max(err, np.abs(dC2.cpu().numpy()[:, :N] - want2).max() / np.abs(want2).max())
    if ret_out:
        return err, dC.cpu().numpy()
    return err


CASES = [
    # M, N, K, layout, epilogue, split_k
    (1000, 256, 256, "kn", EPI_SOFTPLUS, 1),     # forward hidden layer, ragged M
    (512, 213, 43, "kn", EPI_SOFTPLUS, 1),       # first / skip-producing layers: odd K, N not a multiple of 16
    (300, 256, 256, "kn", EPI_BIAS, 1),
    (640, 262, 128, "kk", EPI_ACCUM, 1),         # head dgrad into the shared (feature|x|n) gradient: two n tiles
    (640, 256, 256, "kk", EPI_MUL_S, 1),         # hidden-layer dgrad with the sigmoid factor
    (384, 256, 301, "kn", EPI_ADJ, 1),           # adjoint of the normal pass
    (259, 128, 5000, "mn", EPI_ATOMIC, 7),       # wgrad, contraction over samples, split-K atomics
    (43, 256, 4096, "mn", EPI_ATOMIC, 4),
    (256, 213, 2048, "mn", EPI_ATOMIC, 1),
    (128, 64, 64, "kn", EPI_BIAS, 1),
    (130, 48, 40, "kk", EPI_BIAS, 1),
    # memory-bound corner shapes (gemm_skinny.cu)
    (1000, 1, 256, "kn", EPI_BIAS, 1),           # sdf column
    (1000, 3, 259, "kn", EPI_BIAS, 1),           # colour outputs
    (640, 4, 256, "kk", EPI_ACCUM, 1),           # grid-feature gradient columns
    (333, 6, 128, "kn", EPI_BIAS, 1),
    (512, 256, 1, "kn", EPI_MUL_S, 1),           # rank-1 update with the sigmoid factor
    (512, 128, 6, "kk", EPI_MUL_S, 1),
    (256, 1, 5000, "mn", EPI_ATOMIC, 1),         # weight gradients of the narrow last layers
    (259, 3, 4096, "mn", EPI_ATOMIC, 1),
    (128, 6, 3001, "mn", EPI_ATOMIC, 1),
]


@pytest.mark.parametrize("M,N,K,layout,epi,split_k", CASES)
@pytest.mark.parametrize("tc", [1, 0])
def test_gemm_matches_float64(M, N, K, layout, epi, split_k, tc):
    err = run(M, N, K, layout, epi, tc, split_k)
    print(f"  gemm M={M} N={N} K={K} {layout} epi={epi} tc={tc}: rel err {err:.2e}")
    assert err <= 1e-5, err


PAIR_CASES = [
    (8192, 256, 256, "kn", EPI_SOFTPLUS),      # forward hidden layer on CTA pairs (tcgen05 cta_group::2)
    (4100, 213, 43, "kn", EPI_SOFTPLUS),       # ragged rows, odd K, N rounded up to 256
    (4096, 128, 128, "kn", EPI_BIAS),          # N = 128: 64 columns per CTA
    (6000, 262, 128, "kn", EPI_ACCUM),         # two n tiles
    (5000, 256, 256, "kk", EPI_MUL_S),         # K-major B fallback
    (4608, 256, 301, "kn", EPI_ADJ),
]


@pytest.mark.parametrize("M,N,K,layout,epi", PAIR_CASES)
def test_cta_pair_kernel_matches_float64(M, N, K, layout, epi):
    _lib.call("ndjir_set_option", "mlp_cta_pair", 1)
    try:
        err = run(M, N, K, layout, epi, 1)
    finally:
        _lib.call("ndjir_set_option", "mlp_cta_pair", 0)
    print(f"  pair gemm M={M} N={N} K={K} {layout} epi={epi}: rel err {err:.2e}")
    assert err <= 1e-5, err


def test_tf32_operand_truncation_is_harmless():
    """3xTF32 relies on hi + lo == x with hi = what the tensor core keeps of a raw fp32 operand.  Clearing the low
    mantissa bits explicitly (mlp_mask_hi=1) and leaving them (0) must agree to fp32 accuracy."""
    e0 = run(512, 256, 256, "kn", EPI_BIAS, 1, mask_hi=0)
    e1 = run(512, 256, 256, "kn", EPI_BIAS, 1, mask_hi=1)
    print(f"  raw hi operand: {e0:.2e}; masked hi operand: {e1:.2e}")
    assert e0 <= 1e-5 and e1 <= 1e-5


PRESPLIT_CASES = [c for c in CASES if c[4] != EPI_ATOMIC and c[1] >= 32 and c[2] >= 16]


@pytest.mark.parametrize("M,N,K,layout,epi,split_k", PRESPLIT_CASES)
def test_presplit_weights_are_bit_identical(M, N, K, layout, epi, split_k):
    """A registered B operand (weights) arrives as raw + lo TMA tiles (ndjir_split_lo / ndjir_gemm_presplit) instead
    of being split by the transform warps: same tiles in shared memory, so the product is bit-identical."""
    e0, c0 = run(M, N, K, layout, epi, 1, split_k, presplit=0, ret_out=True)
    e1, c1 = run(M, N, K, layout, epi, 1, split_k, presplit=1, ret_out=True)
    assert e1 <= 1e-5, e1
    assert np.array_equal(c0, c1, equal_nan=True), f"pre-split result differs (errors {e0:.2e} / {e1:.2e})"


@pytest.mark.parametrize("rows,K_in,N,split", [(5000, 256, 256, 7), (4096, 43, 256, 4), (3001, 128, 128, 3),
                                                (2048, 259, 213, 1), (6000, 256, 262, 5), (1000, 128, 6, 1),
                                                (70000, 256, 64, 16)])
@pytest.mark.parametrize("fused", [1, 0])
def test_wgrad_bias_matches_float64(rows, K_in, N, split, fused):
    """ndjir_wgrad_bias: gW += A^T dZ and gb += column sums of dZ in one call (on the tcgen05 path the column sums are
    carried by the transform warps' pass over the dZ tiles) against float64; padding holds NaN."""
    rng = np.random.RandomState(rows + N)
    A = rng.randn(rows, K_in).astype(np.float32)
    dZ = rng.randn(rows, N).astype(np.float32)
    lda, ldz, ldw = r4(K_in) + 4, r4(N) + 4, r4(N) + 8
    As = np.full((rows, lda), np.nan, np.float32); As[:, :K_in] = A
    Zs = np.full((rows, ldz), np.nan, np.float32); Zs[:, :N] = dZ
    gW0 = rng.randn(K_in, ldw).astype(np.float32)
    gb0 = rng.randn(r4(N) + 4).astype(np.float32)
    dA, dZd, gW, gb = dev(As), dev(Zs), dev(gW0), dev(gb0)
    _lib.call("ndjir_set_option", "mlp_fused_colsum", fused)
    try:
        _lib.call("ndjir_wgrad_bias", rows, K_in, N, dA, lda, dZd, ldz, gW, ldw, gb, split, 0)
        torch.cuda.synchronize()
    finally:
        _lib.call("ndjir_set_option", "mlp_fused_colsum", 1)
    wantW = gW0[:, :N].astype(np.float64) + A.astype(np.float64).T @ dZ.astype(np.float64)
    wantb = gb0[:N].astype(np.float64) + dZ.astype(np.float64).sum(axis=0)
    gotW, gotb = gW.cpu().numpy(), gb.cpu().numpy()
    # long contractions per K split: the tensor core's truncating accumulation shows at ~2e-5 of the max (the same with
    # and without the fused column sums); the engine-level gradient bar on this path is 2e-4
    assert np.abs(gotW[:, :N] - wantW).max() / np.abs(wantW).max() <= 5e-5
    assert np.abs(gotb[:N] - wantb).max() / np.abs(wantb).max() <= 1e-5, (gotb[:8], wantb[:8])
    assert np.array_equal(gotW[:, N:], gW0[:, N:]) and np.array_equal(gotb[N:], gb0[N:]), "wrote outside the extents"
