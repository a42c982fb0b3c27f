"""GPU check of the split-fp16 MLP engine (ndjir_gemm_h, csrc/gemm_h.cu + csrc/h16_ops.cu) through the C ABI against
float64: every fused epilogue in both accumulation orders and both output forms, ragged / odd shapes (the 213 + 43
skip layer, 43- and 262-wide gradients, K = 44), the MN-major weight-gradient product with its fused bias-gradient
column sums, the memory-bound corner kernels, pack / unpack round trips and the scale update.

Bars: a product of operands that are EXACTLY representable in the split format must match float64 to 2e-6 with
`precise` (measured 4-6e-7) and 5e-6 without (measured 1.4e-6; the 3xTF32 kernel it replaces: 2-8e-6); the split
representation itself is accurate to 2^-22."""
import numpy as np
import pytest
import torch

from ndjir_b200 import _lib, h16

pytestmark = pytest.mark.gpu


def st():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def softplus(x, beta=100.0):
    return torch.nn.functional.softplus(x, beta=beta, threshold=1e9)


def make(M, N, K, seed=0):
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.rand((M, K), device=dev, generator=g) * 0.5
    W = torch.randn((N, K), device=dev, generator=g) * 0.08
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    Hh = torch.rand((M, N), device=dev, generator=g) * 0.02
    U = torch.randn((M, N), device=dev, generator=g) * 0.3
    sc = h16.Scales(dev)
    b = {k: h16.HBuf(r, c, dev, sc, k) for k, (r, c) in dict(A=(M, K), B=(N, K), H=(M, N), U=(M, N), C=(M, N),
                                                            C2=(M, N)).items()}
    for k, v in dict(A=16.0, B=1024.0, H=64.0, U=8.0, C=32.0, C2=4.0).items():
        sc.scale[b[k].slot] = v
    b["A"].pack(A, st()); b["B"].pack(W, st()); b["H"].pack(Hh, st()); b["U"].pack(U, st())
    vals = {k: b[k].unpack(st()) for k in ("A", "B", "H", "U")}
    assert max(rel(vals["A"], A), rel(vals["B"], W)) < 2.5e-7, "split representation"
    return b, vals, bias, sc, g


SHAPES = [(256, 256, 256), (1000, 256, 256), (4096, 128, 128), (777, 213, 256), (512, 256, 44), (300, 43, 256),
          (640, 262, 128), (40000, 256, 256), (1000, 21, 64), (900, 21, 43), (700, 17, 256)]


@pytest.fixture
def kernel_choice():
    """restores the default kernel selection (resident-weight kernel on) after a test that switches it"""
    yield
    _lib.call("ndjir_set_option", "mlp_h_resident", 0)
    _lib.call("ndjir_set_option", "mlp_h_tma_epi", 1)


@pytest.mark.parametrize("resident", [0, 1, 2])
@pytest.mark.parametrize("precise", [False, True])
@pytest.mark.parametrize("epi", [h16.EPI_BIAS, h16.EPI_SOFTPLUS, h16.EPI_ACCUM, h16.EPI_MUL_S, h16.EPI_ADJ])
def test_products_match_float64(epi, precise, resident, kernel_choice):
    """resident = 0: the streaming kernel (csrc/gemm_h.cu) with the TMA-staged epilogue where the operands allow it (the
    default), 1: the resident-weight kernel (csrc/gemm_h3.cu; K <= 256, always in the `precise` accumulation order),
    2: the streaming kernel with the direct (row-per-lane) epilogue everywhere"""
    _lib.call("ndjir_set_option", "mlp_h_resident", 1 if resident == 1 else 0)
    _lib.call("ndjir_set_option", "mlp_h_tma_epi", 0 if resident == 2 else 1)
    for (M, N, K) in SHAPES:
        for out_h in ((True, False) if (M, N, K) == SHAPES[1] else (True,)):
            b, v, bias, sc, g = make(M, N, K)
            acc = v["A"].double() @ v["B"].double().T
            Cf = torch.zeros((M, N), device="cuda")
            C2f = torch.zeros((M, N), device="cuda")
            kw = dict(A=b["A"].hmat(), B=b["B"].hmat(), precise=precise, bias=bias.data_ptr())
            want2 = None
            if epi == h16.EPI_BIAS:
                want = 0.7 * acc + bias.double(); kw.update(alpha=0.7)
            elif epi == h16.EPI_SOFTPLUS:
                want = 0.5 * softplus(acc + bias.double()); kw.update(out_scale=0.5)
            elif epi == h16.EPI_ACCUM:
                Cf = torch.randn((M, N), device="cuda", generator=g)
                want = Cf.double() + 0.7 * acc; kw.update(alpha=0.7); out_h = False
            elif epi == h16.EPI_MUL_S:
                s = 1.0 - torch.exp(-100.0 * v["H"].double())
                want = 0.9 * acc * s + v["U"].double()
                kw.update(alpha=0.9, Hh=b["H"].hmat(), Uh=b["U"].hmat())
            else:
                s = 1.0 - torch.exp(-100.0 * v["H"].double())
                want, want2 = acc * v["U"].double() * 100.0 * (1 - s), 0.5 * acc * s
                kw.update(out_scale=0.5, Hh=b["H"].hmat(), Uh=b["U"].hmat())
                kw.update(C2h=b["C2"].hmat()) if out_h else kw.update(C2=C2f.data_ptr(), ldc2=N)
            kw.update(Ch=b["C"].hmat()) if out_h else kw.update(C=Cf.data_ptr(), ldc=N)
            h16.gemm_h(st(), M, N, K, epi, **kw)
            torch.cuda.synchronize()
            tol = 2e-6 if (precise or (resident == 1 and K <= 256)) else 5e-6
            got = b["C"].unpack(st()) if out_h else Cf
            assert rel(got, want) < tol, (M, N, K, epi, precise, out_h, rel(got, want))
            if want2 is not None:
                got2 = b["C2"].unpack(st()) if out_h else C2f
                assert rel(got2, want2) < tol, (M, N, K, "C2", rel(got2, want2))
            if out_h:     # the running maximum of the written tensor (what the next step's scale is derived from)
                assert abs(float(sc.amax[b["C"].slot]) - float(want.abs().max())) <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize("tall", [0, 1])
@pytest.mark.parametrize("rows,Kin,N,split", [(4096, 256, 256, 4), (10000, 256, 213, 8), (5000, 44, 256, 3),
                                              (8192, 262, 128, 16), (65536, 256, 256, 148), (7000, 200, 256, 5)])
def test_weight_gradient_with_fused_bias_gradient(rows, Kin, N, split, tall):
    """tall = 1: work items of 256 rows of A^T (two accumulators, the dZ tile staged once; measured slower, off by
    default); 0: 128 rows"""
    _lib.call("ndjir_set_option", "mlp_h_tall", tall)
    try:
        _weight_gradient_case(rows, Kin, N, split)
    finally:
        _lib.call("ndjir_set_option", "mlp_h_tall", 0)


def _weight_gradient_case(rows, Kin, N, split):
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand((rows, Kin), device=dev, generator=g) * 0.5
    dZ = torch.randn((rows, N), device=dev, generator=g) * 1e-5
    sc = h16.Scales(dev)
    bA, bZ = h16.HBuf(rows, Kin, dev, sc, "A"), h16.HBuf(rows, N, dev, sc, "Z")
    sc.scale[bA.slot] = 16.0; sc.scale[bZ.slot] = 2.0 ** 24
    bA.pack(A, st()); bZ.pack(dZ, st())
    A_, Z_ = bA.unpack(st()), bZ.unpack(st())
    gW = torch.zeros((Kin, N), device=dev)
    gb = torch.zeros(N, device=dev)
    h16.gemm_h(st(), Kin, N, rows, h16.EPI_ATOMIC, A=bA.hmat(), B=bZ.hmat(), mn_major=True, split_k=split,
               C=gW.data_ptr(), ldc=N, colsum=gb.data_ptr())
    cs = torch.zeros(N, device=dev)
    _lib.call("ndjir_colsum_h", rows, N, cs, bZ.hmat(track=False), 1.0, st())
    torch.cuda.synchronize()
    assert rel(gW, A_.double().T @ Z_.double()) < 1e-5
    assert rel(gb, Z_.double().sum(0)) < 2e-6 and rel(cs, Z_.double().sum(0)) < 2e-6


def test_corner_kernels_match_float64():
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(2)
    M, K = 5000, 256
    A = torch.rand((M, K), device=dev, generator=g)
    sc = h16.Scales(dev)
    bA = h16.HBuf(M, K, dev, sc, "A")
    sc.scale[bA.slot] = 16.0
    bA.pack(A, st())
    A_ = bA.unpack(st())
    for N in (1, 3, 6):
        W = torch.randn((K, N), device=dev, generator=g) * 0.1
        b = torch.randn(8, device=dev, generator=g)
        C = torch.zeros((M, 16), device=dev)
        h16.gemm_h(st(), M, N, K, h16.EPI_BIAS, A=bA.hmat(), B32=W.data_ptr(), b_rs=N, b_cs=1, C=C.data_ptr(), ldc=16,
                   bias=b.data_ptr())
        assert rel(C[:, :N], A_.double() @ W.double() + b[:N].double()) < 1e-6
        dZ = torch.randn((M, 16), device=dev, generator=g)
        gW = torch.zeros((K, N), device=dev)
        h16.gemm_h(st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), mn_major=True, B32=dZ.data_ptr(), b_rs=16, b_cs=1,
                   C=gW.data_ptr(), ldc=N)
        assert rel(gW, A_.double().T @ dZ[:, :N].double()) < 2e-6
    for N, Kk in ((256, 1), (256, 3), (128, 6), (213, 2)):
        X = torch.randn((M, 16), device=dev, generator=g)
        W = torch.randn((Kk, N), device=dev, generator=g) * 0.1
        Hh = torch.rand((M, N), device=dev, generator=g) * 0.02
        bH, bC = h16.HBuf(M, N, dev, sc, f"H{N}"), h16.HBuf(M, N, dev, sc, f"C{N}")
        sc.scale[bH.slot] = 64.0; sc.scale[bC.slot] = 128.0
        bH.pack(Hh, st())
        H_ = bH.unpack(st())
        h16.gemm_h(st(), M, N, Kk, h16.EPI_MUL_S, A32=X.data_ptr(), a_rs=16, a_cs=1, B32=W.data_ptr(), b_rs=N, b_cs=1,
                   Ch=bC.hmat(), Hh=bH.hmat())
        want = (X[:, :Kk].double() @ W.double()) * (1 - torch.exp(-100.0 * H_.double()))
        assert rel(bC.unpack(st()), want) < 1e-6, (N, Kk)


def test_pack_scale_update_round_trip():
    """delayed scaling: pack records max|x|; ndjir_scale_update turns it into the power of two that puts the largest
    value in [2^9, 2^10); values far below the maximum keep 22 bits, clamping is flagged."""
    dev = torch.device("cuda")
    sc = h16.Scales(dev)
    x = torch.randn((1000, 96), device=dev) * 3e-7
    b = h16.HBuf(1000, 96, dev, sc, "x", init=2.0 ** 24)
    b.pack(x, st())
    sc.update(st())
    torch.cuda.synchronize()
    s = float(sc.scale[b.slot])
    amax = float(x.abs().max())
    assert 2 ** 9 <= s * amax < 2 ** 10 and s == 2.0 ** round(np.log2(s))
    b.pack(x, st())
    assert rel(b.unpack(st()), x) < 2.5e-7
    # a tensor that outgrew its scale is clamped and flagged at the next update
    b.pack(x * 1000.0, st())
    sc.update(st())
    torch.cuda.synchronize()
    assert int(sc.flags[0]) & 1
    assert torch.isfinite(b.unpack(st())).all()


@pytest.mark.parametrize("cols,col", [(262, 0), (43, 8), (9, 64), (6, 16), (301, 0), (13, 3)])
def test_pack_ragged_width_into_a_view(cols, col):
    """ndjir_pack_h of a width that is no multiple of 8 into columns [col, col + cols) of a wider split matrix: one
    launch (16-byte groups + a ragged last group); bit-identical to packing column by column, and the columns either
    side of the view keep their contents."""
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(cols)
    rows, wide = 777, col + cols + 40
    src = torch.randn((rows, cols + 2), device=dev, generator=g)[:, :cols]        # row stride cols + 2
    full = torch.randn((rows, wide), device=dev, generator=g)
    a, b, keep = h16.HBuf(rows, wide, dev), h16.HBuf(rows, wide, dev), h16.HBuf(rows, wide, dev)
    for buf in (a, b, keep):
        buf.pack(full, st())
    a.pack(src, st(), cols=cols, col=col)
    for c in range(cols):                                                          # 1-wide packs: the scalar path
        _lib.call("ndjir_pack_h", rows, 1, src.data_ptr() + 4 * c, src.stride(0), 1, 1.0, b.hmat(col + c), st())
    torch.cuda.synchronize()
    assert torch.equal(a.t, b.t)
    assert torch.equal(a.t[:, :, :col], keep.t[:, :, :col]) and torch.equal(a.t[:, :, col + cols:], keep.t[:, :, col + cols:])
    assert rel(a.unpack(st())[:, col:col + cols], src) < 2.5e-7


@pytest.mark.parametrize("M,N,K,ldc", [(1000, 70, 32, 72), (600, 43, 256, 44), (900, 256, 128, 264), (5000, 21, 64, 24)])
def test_accumulate_into_fp32_view_with_ragged_width(M, N, K, ldc):
    """C (fp32 rows, row stride ldc > N: a view into a wider matrix) += alpha * A B^T through the TMA-staged epilogue:
    the columns of the parent matrix beyond N must stay untouched, the last 16-column chunk is partial."""
    b, v, bias, sc, g = make(M, N, K)
    C = torch.randn((M, ldc), device="cuda", generator=g)
    want = C.clone().double()
    want[:, :N] += 0.7 * (v["A"].double() @ v["B"].double().T)
    h16.gemm_h(st(), M, N, K, h16.EPI_ACCUM, A=b["A"].hmat(), B=b["B"].hmat(), alpha=0.7, C=C.data_ptr(), ldc=ldc)
    torch.cuda.synchronize()
    assert rel(C[:, :N], want[:, :N]) < 5e-6
    assert torch.equal(C[:, N:].double(), want[:, N:])
