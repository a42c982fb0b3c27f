"""Grid-feature micro-benchmark (BASELINE.json config 3; shapes from the reference's scripts/bench_voxel_hash.py:29-45,
79-84 and scripts/bench_lanczos_voxel.py:27-39,73-75, plus the training shapes).  Times our C-ABI kernels and, when
oracle/_ref is present, the reference kernels recompiled for sm_100a on the same inputs.  CUDA events on stream 0,
3 warm-ups, L2 flushed between timed iterations, 10 iterations like the reference scripts.

  python tools/bench_grid.py [--log2-points 24] [--out gpurun_out/bench_grid.json] [--only voxel,hash,...]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndjir_b200 import compat  # noqa: E402
from ndjir_b200._lib import call  # noqa: E402

MN, MX = [-1.0] * 3, [1.0] * 3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class Timer:
    def __init__(self):
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 256 MiB > 126 MB L2

    def time(self, fn, iters=10, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            self.flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.mean(ts)), float(np.min(ts))


def ref_or_none(name):
    try:
        from oracle import build_ref
        return build_ref.load(name)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-points", type=int, default=24)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_grid.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    B = 1 << args.log2_points
    hbm, which = peaks()
    T = Timer()
    rng = np.random.RandomState(412)
    q = torch.as_tensor((rng.rand(B, 3) * 2 - 1).astype(np.float32)).cuda()
    results = []

    def record(name, kind, ms_mean, ms_min, bytes_per_pt, impl, extra=None):
        gbs = bytes_per_pt * B / (ms_mean * 1e-3) / 1e9
        r = dict(kernel=name, pass_=kind, impl=impl, points=B, ms=ms_mean, ms_min=ms_min,
                 algorithmic_bytes_per_point=bytes_per_pt, GBps=gbs, frac_of_hbm_peak=gbs / hbm, peak_GBps=hbm,
                 peak_source=which)
        if extra:
            r.update(extra)
        results.append(r)
        print(json.dumps(r), flush=True)

    def run_family(tag, modname, fwd_name, feat, spec, D, C, bytes_fwd, bytes_bwd, hash_args=None, agg_modes=(0,),
                   ours_label="ndjir_b200", with_ref=True):
        if only and tag not in only:
            return
        ours, ref = compat.load(modname), (None if (args.no_ref or not with_ref) else ref_or_none(modname))
        out = torch.empty(B * C, device="cuda")
        go = torch.ones(B * C, device="cuda")
        gf = torch.zeros_like(feat)
        N = B * C if hash_args is None else B * hash_args[3]
        for impl, mod in ((ours_label, ours), ("reference_sm100a", ref)):
            if mod is None:
                continue
            if hash_args is None:
                fwd = lambda: getattr(mod, fwd_name)(N, out.data_ptr(), q.data_ptr(), feat.data_ptr(), spec, D, MN, MX, False)
                bwd = lambda: mod.grad_feature(N, gf.data_ptr(), go.data_ptr(), q.data_ptr(), spec, D, MN, MX, False, True)
            else:
                fwd = lambda: mod.voxel_hash_feature(N, out.data_ptr(), q.data_ptr(), feat.data_ptr(), *hash_args, MN, MX, False)
                bwd = lambda: mod.grad_feature(N, gf.data_ptr(), go.data_ptr(), q.data_ptr(), *hash_args, MN, MX, False, True)
            m, mi = T.time(fwd, args.iters)
            record(tag, "fwd", m, mi, bytes_fwd, impl)
            for agg in (agg_modes if impl == ours_label else (0,)):
                if impl == ours_label:
                    call("ndjir_set_option", "scatter_aggregate", agg)
                m, mi = T.time(bwd, args.iters)
                record(tag, "grad_feature", m, mi, bytes_bwd, impl, dict(scatter_aggregate=agg))
            call("ndjir_set_option", "scatter_aggregate", 0)
        del out, go, gf

    # trilinear voxel at the training shape: G=512, D=4 (2 GiB table)
    if not only or "voxel" in only:
        G, D = 512, 4
        feat = (torch.randn(G, G, G, D, device="cuda") * 0.01)
        # default dispatch: brick-ordered sweep (voxel_binned.cu) for >= 2^21 points on a table >= 96 MB
        run_family("voxel", "voxel_feature_cuda", "query_on_voxel", feat, (G, G, G), D, D, 156, 284,
                   ours_label="ndjir_b200 (brick-ordered, default)", with_ref=False)
        call("ndjir_set_option", "voxel_binned", 0)
        run_family("voxel", "voxel_feature_cuda", "query_on_voxel", feat, (G, G, G), D, D, 156, 284, agg_modes=(0, 1),
                   ours_label="ndjir_b200 (direct)")
        call("ndjir_set_option", "voxel_binned", -1)
        del feat
    # triplane / triline at triplaneline.yaml: G=2048, D=8
    if not only or "triplane" in only:
        G, D = 2048, 8
        feat = torch.randn(3, G, G, D, device="cuda") * 0.01
        run_family("triplane", "triplane_feature_cuda", "query_on_triplane", feat, G, D, 3 * D, 492, 12 + 96 + 2 * 384)
        del feat
    if not only or "triline" in only:
        G, D = 2048, 8
        feat = torch.randn(3, G, D, device="cuda") * 0.01
        run_family("triline", "triline_feature_cuda", "query_on_triline", feat, G, D, 3 * D, 108, 108)
        del feat
    # voxel hash bench shape: G0=16, gf=1.5, T0=2^15, L=16, D=2 (3.8 MB table, L2 resident)
    if not only or "hash" in only:
        G0, gfac, T0, L, D = 16, 1.5, 2 ** 15, 16, 2
        n = call("ndjir_voxel_hash_num_params", G0, gfac, T0, L, D)
        feat = torch.randn(n, device="cuda") * 0.01
        run_family("hash", "voxel_hash_feature_cuda", None, feat, None, D, D * L, 140, 140 + 0, hash_args=(G0, gfac, T0, L, D))
        del feat
    # Lanczos voxel bench shape: G=256, D=4
    if not only or "lanczos" in only:
        G, D = 256, 4
        feat = torch.randn(G, G, G, D, device="cuda") * 0.01
        # default dispatch: brick-ordered (268 MB table, 2^24 points); then the direct kernels
        run_family("lanczos", "lanczos_voxel_feature_cuda", "query_on_voxel", feat, (G, G, G), D, D, 44, 44 + 16,
                   ours_label="ndjir_b200 (brick-ordered, default)", with_ref=False)
        call("ndjir_set_option", "voxel_binned", 0)
        run_family("lanczos", "lanczos_voxel_feature_cuda", "query_on_voxel", feat, (G, G, G), D, D, 44, 44 + 16,
                   ours_label="ndjir_b200 (direct)")
        call("ndjir_set_option", "voxel_binned", -1)
        del feat
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
