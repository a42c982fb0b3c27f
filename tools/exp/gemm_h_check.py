"""Bring-up check + micro-benchmark of the split-fp16 tcgen05 product kernel (run on the GPU box):
python tools/exp/gemm_h_check.py [--bench]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ndjir_b200 import _lib, h16  # noqa: E402

dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def rel(a, b):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def softplus(x, beta=100.0):
    return torch.nn.functional.softplus(x, beta=beta, threshold=1e9)


def check_kk(M, N, K, epi, precise, sa=16.0, sb=1024.0, out_h=True, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.rand((M, K), device=dev, generator=g) * 0.5
    W = torch.randn((N, K), device=dev, generator=g) * 0.08           # B as (N x K) rows = W^T
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    Hh = torch.rand((M, N), device=dev, generator=g) * 0.02
    U = torch.randn((M, N), device=dev, generator=g) * 0.3
    sc = h16.Scales(dev)
    bA = h16.HBuf(M, K, dev, sc, "A"); bB = h16.HBuf(N, K, dev, sc, "B"); bH = h16.HBuf(M, N, dev, sc, "H")
    bU = h16.HBuf(M, N, dev, sc, "U"); bC = h16.HBuf(M, N, dev, sc, "C"); bC2 = h16.HBuf(M, N, dev, sc, "C2")
    sc.scale[bA.slot] = sa; sc.scale[bB.slot] = sb; sc.scale[bH.slot] = 64.0; sc.scale[bU.slot] = 8.0
    sc.scale[bC.slot] = 32.0; sc.scale[bC2.slot] = 4.0
    bA.pack(A, st()); bB.pack(W, st()); bH.pack(Hh, st()); bU.pack(U, st())
    A_, W_, H_, U_ = bA.unpack(st()), bB.unpack(st()), bH.unpack(st()), bU.unpack(st())
    e_rep = max(rel(A_, A), rel(W_, W))
    acc = (A_.double() @ W_.double().T)
    Cf = torch.zeros((M, N), device=dev)
    C2f = torch.zeros((M, N), device=dev)
    kw = dict(A=bA.hmat(), B=bB.hmat(), precise=precise, bias=bias.data_ptr())
    if epi == h16.EPI_BIAS:
        want = 0.7 * acc + bias.double()
        kw.update(alpha=0.7)
    elif epi == h16.EPI_SOFTPLUS:
        want = 0.5 * softplus(acc + bias.double())
        kw.update(out_scale=0.5)
    elif epi == h16.EPI_ACCUM:
        Cf = torch.randn((M, N), device=dev, generator=g)
        want = Cf.double() + 0.7 * acc
        kw.update(alpha=0.7)
        out_h = False
    elif epi == h16.EPI_MUL_S:
        s = 1.0 - torch.exp(-100.0 * H_.double())
        want = 0.9 * acc * s + U_.double()
        kw.update(alpha=0.9, Hh=bH.hmat(), Uh=bU.hmat())
    elif epi == h16.EPI_ADJ:
        s = 1.0 - torch.exp(-100.0 * H_.double())
        want = acc * U_.double() * 100.0 * (1 - s)
        want2 = 0.5 * acc * s
        kw.update(out_scale=0.5, Hh=bH.hmat(), Uh=bU.hmat())
        if out_h:
            kw.update(C2h=bC2.hmat())
        else:
            kw.update(C2=C2f.data_ptr(), ldc2=N)
    if out_h:
        kw.update(Ch=bC.hmat())
    else:
        kw.update(C=Cf.data_ptr(), ldc=N)
    h16.gemm_h(st(), M, N, K, epi, **kw)
    torch.cuda.synchronize()
    got = bC.unpack(st()) if out_h else Cf
    e = rel(got, want)
    e2 = None
    if epi == h16.EPI_ADJ:
        got2 = bC2.unpack(st()) if out_h else C2f
        e2 = rel(got2, want2)
    am = float(sc.amax[bC.slot]) if out_h else None
    print(f"KK M={M} N={N} K={K} epi={epi} precise={int(precise)} out_h={int(out_h)}: err {e:.2e}"
          + (f" err2 {e2:.2e}" if e2 is not None else "") + f"  (repr {e_rep:.1e}, amax {am}, want max {float(want.abs().max()):.3g})")
    return e


def check_mn(rows, Kin, N, split, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.rand((rows, Kin), device=dev, generator=g) * 0.5
    dZ = torch.randn((rows, N), device=dev, generator=g) * 1e-5
    sc = h16.Scales(dev)
    bA = h16.HBuf(rows, Kin, dev, sc, "A"); bZ = h16.HBuf(rows, N, dev, sc, "Z")
    sc.scale[bA.slot] = 16.0; sc.scale[bZ.slot] = 2.0 ** 24
    bA.pack(A, st()); bZ.pack(dZ, st())
    A_, Z_ = bA.unpack(st()), bZ.unpack(st())
    want = A_.double().T @ Z_.double()
    gW = torch.zeros((Kin, N), device=dev)
    cf = torch.zeros(N, device=dev)
    h16.gemm_h(st(), Kin, N, rows, h16.EPI_ATOMIC, A=bA.hmat(), B=bZ.hmat(), mn_major=True, split_k=split,
               C=gW.data_ptr(), ldc=N, colsum=cf.data_ptr())
    cs = torch.zeros(N, device=dev)
    _lib.call("ndjir_colsum_h", rows, N, cs, bZ.hmat(track=False), 1.0, st())
    torch.cuda.synchronize()
    e = rel(gW, want)
    ec = rel(cs, Z_.double().sum(0))
    ef = rel(cf, Z_.double().sum(0))
    print(f"MN rows={rows} Kin={Kin} N={N} split={split}: err {e:.2e} colsum {ec:.2e} fused colsum {ef:.2e} (repr {rel(Z_, dZ):.1e})")
    e = max(e, ef)
    return e


def check_corner(seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
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
        want = A_.double() @ W.double() + b[:N].double()
        print(f"skinny_n N={N}: err {rel(C[:, :N], want):.2e}")
        # weight gradient of the same layer
        dZ = torch.randn((M, 16), device=dev, generator=g)
        gW = torch.zeros((K, N), device=dev)
        h16.gemm_h(st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), mn_major=True, B32=dZ.data_ptr(), b_rs=16, b_cs=1,
                   C=gW.data_ptr(), ldc=N)
        want = A_.double().T @ dZ[:, :N].double()
        print(f"skinny_w N={N}: err {rel(gW, want):.2e}")
    # rank-K update with the fused epilogue
    N = 256
    for Kk in (1, 3):
        X = torch.randn((M, 16), device=dev, generator=g)
        W = torch.randn((Kk, N), device=dev, generator=g) * 0.1
        Hh = torch.rand((M, N), device=dev, generator=g) * 0.02
        bH = h16.HBuf(M, N, dev, sc, "H"); bC = h16.HBuf(M, N, dev, sc, "C")
        sc.scale[bH.slot] = 64.0; sc.scale[bC.slot] = 128.0
        bH.pack(Hh, st())
        H_ = bH.unpack(st())
        h16.gemm_h(st(), M, N, Kk, h16.EPI_MUL_S, A32=X.data_ptr(), a_rs=16, a_cs=1, B32=W.data_ptr(), b_rs=N, b_cs=1,
                   Ch=bC.hmat(), Hh=bH.hmat())
        want = (X[:, :Kk].double() @ W.double()) * (1 - torch.exp(-100.0 * H_.double()))
        print(f"skinny_k K={Kk}: err {rel(bC.unpack(st()), want):.2e}")


def bench(M=262144, N=256, K=256, iters=20):
    sc = h16.Scales(dev)
    bA = h16.HBuf(M, K, dev, sc, "A"); bB = h16.HBuf(N, K, dev, sc, "B"); bH = h16.HBuf(M, N, dev, sc, "H")
    bU = h16.HBuf(M, N, dev, sc, "U"); bC = h16.HBuf(M, N, dev, sc, "C"); bC2 = h16.HBuf(M, N, dev, sc, "C2")
    for b in (bA, bB, bH, bU):
        b.t.normal_(0, 0.1)
    bias = torch.zeros(N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gW = torch.zeros((K, N), device=dev)

    def run(name, fn, flops):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts))
        print(f"{name:<40s} {t:.4f} ms  {flops / t / 1e9:.1f} TFLOP/s")

    fl = 2.0 * M * N * K
    _lib.call("ndjir_set_option", "mlp_h_resident", 0)
    for pair in (0, 1):     # 1: TMA-staged epilogue, 0: direct row-per-lane epilogue (reported as "resident=" for history)
        _lib.call("ndjir_set_option", "mlp_h_tma_epi", pair)
        for precise in (0, 1):
            run(f"fwd softplus precise={precise} tma_epi={pair}", lambda: h16.gemm_h(
                st(), M, N, K, h16.EPI_SOFTPLUS, A=bA.hmat(), B=bB.hmat(), precise=precise, Ch=bC.hmat(),
                bias=bias.data_ptr()), fl)
        run(f"dgrad mul_s tma_epi={pair}", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_MUL_S, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), Hh=bH.hmat()), fl)
        run(f"dgrad mul_s+U tma_epi={pair}", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_MUL_S, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), Hh=bH.hmat(), Uh=bU.hmat()), fl)
        run(f"adjoint tma_epi={pair}", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), C2h=bC2.hmat(), Hh=bH.hmat(),
            Uh=bU.hmat()), fl)
    _lib.call("ndjir_set_option", "mlp_h_resident", 0)
    _lib.call("ndjir_set_option", "mlp_h_pair", 1)
    for dbg in ():
        _lib.call("ndjir_set_option", "mlp_h_dbg", dbg)
        for precise in (0, 1):
            run(f"fwd softplus precise={precise} dbg={dbg} PAIR", lambda: h16.gemm_h(
                st(), M, N, K, h16.EPI_SOFTPLUS, A=bA.hmat(), B=bB.hmat(), precise=precise, Ch=bC.hmat(),
                bias=bias.data_ptr()), fl)
        run(f"adjoint dbg={dbg} PAIR", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), C2h=bC2.hmat(), Hh=bH.hmat(),
            Uh=bU.hmat()), fl)
    _lib.call("ndjir_set_option", "mlp_h_pair", 0)
    _lib.call("ndjir_set_option", "mlp_h_dbg", 0)
    for pf in ():
        for dbg in (0, 1, 3):
            _lib.call("ndjir_set_option", "mlp_h_dbg", dbg | (pf << 4))
            run(f"fwd softplus dbg={dbg} resident=1 prefetch={pf}", lambda: h16.gemm_h(
                st(), M, N, K, h16.EPI_SOFTPLUS, A=bA.hmat(), B=bB.hmat(), precise=1, Ch=bC.hmat(),
                bias=bias.data_ptr()), fl)
        _lib.call("ndjir_set_option", "mlp_h_dbg", pf << 4)
        run(f"adjoint resident=1 prefetch={pf}", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), C2h=bC2.hmat(), Hh=bH.hmat(),
            Uh=bU.hmat()), fl)
    for res in (1, 0):
        _lib.call("ndjir_set_option", "mlp_h_resident", res)
        for dbg in (1, 2, 3):
            _lib.call("ndjir_set_option", "mlp_h_dbg", dbg)
            for precise in ((0, 1) if not res else (1,)):
                run(f"fwd softplus precise={precise} dbg={dbg} resident={res}", lambda: h16.gemm_h(
                    st(), M, N, K, h16.EPI_SOFTPLUS, A=bA.hmat(), B=bB.hmat(), precise=precise, Ch=bC.hmat(),
                    bias=bias.data_ptr()), fl)
            if dbg == 1:
                run(f"adjoint dbg={dbg} resident={res}", lambda: h16.gemm_h(
                    st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), C2h=bC2.hmat(), Hh=bH.hmat(),
                    Uh=bU.hmat()), fl)
    _lib.call("ndjir_set_option", "mlp_h_dbg", 0)
    _lib.call("ndjir_set_option", "mlp_h_resident", 0)
    if "--short" in sys.argv and "--wgrad" not in sys.argv:
        return
    for precise in (0, 1):
        run(f"dgrad mul_s+U precise={precise}", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_MUL_S, A=bA.hmat(), B=bB.hmat(), precise=precise, Ch=bC.hmat(), Hh=bH.hmat(),
            Uh=bU.hmat()), fl)
        run(f"adjoint precise={precise}", lambda: h16.gemm_h(
            st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), precise=precise, Ch=bC.hmat(), C2h=bC2.hmat(),
            Hh=bH.hmat(), Uh=bU.hmat()), fl)
    cs0 = torch.zeros(N, device=dev)
    if "--wgrad" in sys.argv:
        for tall in (0, 1):
            _lib.call("ndjir_set_option", "mlp_h_tall", tall)
            for split in (74, 148, 296):
                run(f"wgrad tall={tall} split={split}", lambda: h16.gemm_h(
                    st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), B=bC.hmat(), mn_major=True, split_k=split, C=gW.data_ptr(),
                    ldc=N), fl)
                run(f"wgrad + fused colsum tall={tall} split={split}", lambda: h16.gemm_h(
                    st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), B=bC.hmat(), mn_major=True, split_k=split, C=gW.data_ptr(),
                    ldc=N, colsum=cs0.data_ptr()), fl)
        _lib.call("ndjir_set_option", "mlp_h_tall", 0)
        return
    for split in (148, 296, 592):
        run(f"wgrad split={split}", lambda: h16.gemm_h(
            st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), B=bC.hmat(), mn_major=True, split_k=split, C=gW.data_ptr(),
            ldc=N), fl)
        run(f"wgrad + fused colsum split={split}", lambda: h16.gemm_h(
            st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), B=bC.hmat(), mn_major=True, split_k=split, C=gW.data_ptr(),
            ldc=N, colsum=cs0.data_ptr()), fl)
    cs = torch.zeros(N, device=dev)
    run("colsum_h", lambda: _lib.call("ndjir_colsum_h", M, N, cs, bC.hmat(track=False), 1.0, st()), fl)
    src = torch.randn((M, N), device=dev)
    run("pack_h", lambda: bC.pack(src, st()), fl)


def ncu_once(M=262144, N=256, K=256):
    """one launch of every hidden-layer product type (for `ncu --set full -k regex:gemm_h_kernel`)"""
    sc = h16.Scales(dev)
    bA = h16.HBuf(M, K, dev, sc, "A"); bB = h16.HBuf(N, K, dev, sc, "B"); bH = h16.HBuf(M, N, dev, sc, "H")
    bU = h16.HBuf(M, N, dev, sc, "U"); bC = h16.HBuf(M, N, dev, sc, "C"); bC2 = h16.HBuf(M, N, dev, sc, "C2")
    for b in (bA, bB, bH, bU):
        b.t.normal_(0, 0.1)
    bias = torch.zeros(N, device=dev)
    gW = torch.zeros((K, N), device=dev)
    cs0 = torch.zeros(N, device=dev)
    h16.gemm_h(st(), M, N, K, h16.EPI_SOFTPLUS, A=bA.hmat(), B=bB.hmat(), precise=1, Ch=bC.hmat(), bias=bias.data_ptr())
    h16.gemm_h(st(), M, N, K, h16.EPI_MUL_S, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), Hh=bH.hmat())
    h16.gemm_h(st(), M, N, K, h16.EPI_MUL_S, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), Hh=bH.hmat(), Uh=bU.hmat())
    h16.gemm_h(st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), C2h=bC2.hmat(), Hh=bH.hmat(),
               Uh=bU.hmat())
    h16.gemm_h(st(), K, N, M, h16.EPI_ATOMIC, A=bA.hmat(), B=bC.hmat(), mn_major=True, split_k=148, C=gW.data_ptr(),
               ldc=N, colsum=cs0.data_ptr())
    torch.cuda.synchronize()


def ncu_resident(M=262144, N=256, K=256):
    """resident-weight kernel: forward, forward without epilogue traffic and with one product per K step, adjoint"""
    sc = h16.Scales(dev)
    bA = h16.HBuf(M, K, dev, sc, "A"); bB = h16.HBuf(N, K, dev, sc, "B"); bH = h16.HBuf(M, N, dev, sc, "H")
    bU = h16.HBuf(M, N, dev, sc, "U"); bC = h16.HBuf(M, N, dev, sc, "C"); bC2 = h16.HBuf(M, N, dev, sc, "C2")
    for b in (bA, bB, bH, bU):
        b.t.normal_(0, 0.1)
    bias = torch.zeros(N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for dbg in (0, 3, 1):
        _lib.call("ndjir_set_option", "mlp_h_dbg", dbg)
        flush.zero_()
        h16.gemm_h(st(), M, N, K, h16.EPI_SOFTPLUS, A=bA.hmat(), B=bB.hmat(), precise=1, Ch=bC.hmat(), bias=bias.data_ptr())
    _lib.call("ndjir_set_option", "mlp_h_dbg", 0)
    flush.zero_()
    h16.gemm_h(st(), M, N, K, h16.EPI_ADJ, A=bA.hmat(), B=bB.hmat(), Ch=bC.hmat(), C2h=bC2.hmat(), Hh=bH.hmat(),
               Uh=bU.hmat())
    torch.cuda.synchronize()


if __name__ == "__main__":
    if "--ncu-res" in sys.argv:
        ncu_resident()
        sys.exit(0)
    if "--ncu" in sys.argv:
        ncu_once()
        sys.exit(0)
    if "--bench" in sys.argv:
        bench()
        sys.exit(0)
    bad = 0
    for epi in (h16.EPI_BIAS, h16.EPI_SOFTPLUS, h16.EPI_ACCUM, h16.EPI_MUL_S, h16.EPI_ADJ):
        for precise in (False, True):
            for (M, N, K) in ((256, 256, 256), (1000, 256, 256), (4096, 128, 128), (777, 213, 256), (512, 256, 44),
                              (300, 43, 256), (640, 262, 128), (40000, 256, 256)):
                e = check_kk(M, N, K, epi, precise)
                bad += e > 2e-5
    e = check_kk(2048, 256, 256, h16.EPI_SOFTPLUS, True, out_h=False)
    e = check_kk(2048, 256, 256, h16.EPI_ADJ, True, out_h=False)
    for (rows, Kin, N, split) in ((4096, 256, 256, 4), (10000, 256, 213, 8), (5000, 44, 256, 3), (8192, 262, 128, 16),
                                  (65536, 256, 256, 148)):
        e = check_mn(rows, Kin, N, split)
        bad += e > 2e-5
    check_corner()
    print("BAD" if bad else "ALL OK", bad)
