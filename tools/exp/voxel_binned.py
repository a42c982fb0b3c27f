"""Experiment: brick-ordered voxel gather / scatter vs the direct kernels at the micro-benchmark shape
(2^24 uniform points, 512^3 x 4 table).  CUDA events, L2 flushed between iterations.
  python tools/exp/voxel_binned.py [--log2-points 24] [--mbs 4,8,16,32,64]"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndjir_b200._lib import call  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2-points", type=int, default=24)
ap.add_argument("--mbs", default="4,8,16,32,64")
ap.add_argument("--G", type=int, default=512)
ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (32/64/128), 0 = leave")
args = ap.parse_args()
if args.l2_fetch:
    import ctypes
    torch.cuda.init(); torch.zeros(1, device="cuda")
    rt = ctypes.CDLL("libcudart.so.12")
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print("L2 fetch granularity before:", v.value)
    print("set ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(args.l2_fetch)))
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print("L2 fetch granularity after:", v.value)
B, G, D = 1 << args.log2_points, args.G, 4
MN, MX = [-1.0] * 3, [1.0] * 3
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
rng = np.random.RandomState(412)
q = torch.as_tensor((rng.rand(B, 3) * 2 - 1).astype(np.float32)).cuda()
feat = torch.randn(G, G, G, D, device="cuda") * 0.01
out = torch.empty(B, D, device="cuda")
go = torch.ones(B, D, device="cuda")
gf = torch.zeros_like(feat)
wsb = call("ndjir_voxel_binned_workspace_bytes", B)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.mean(ts))


def show(name, ms, bpp):
    gbs = bpp * B / ms / 1e6
    print(f"{name:56s} {ms:8.3f} ms  {gbs:8.1f} GB/s algorithmic ({gbs / hbm:.3f} of {hbm})", flush=True)


call("ndjir_set_option", "voxel_binned", 0)
show("gather direct (gather4)", t(lambda: call("ndjir_voxel_query_on_voxel", B, out, q, feat, [G] * 3, D, MN, MX, 0, 0)), 156)
ref_out = out.clone()
show("scatter direct accum=0 (zero-fill + scatter8)", t(lambda: call("ndjir_voxel_grad_feature", B, gf, go, q, [G] * 3, D, MN, MX, 0, 0)), 284)
show("scatter direct accum=1", t(lambda: call("ndjir_voxel_grad_feature", B, gf, go, q, [G] * 3, D, MN, MX, 1, 0)), 284)
call("ndjir_voxel_grad_feature", B, gf, go, q, [G] * 3, D, MN, MX, 0, 0)
ref_gf = gf.clone()
for mb in [int(x) for x in args.mbs.split(",")]:
    call("ndjir_set_option", "voxel_bin_mb", mb)
    show(f"gather binned, {mb} MiB bricks (explicit workspace)",
         t(lambda: call("ndjir_voxel_query_on_voxel_binned", B, out, q, feat, [G] * 3, D, MN, MX, 0, ws, wsb, 0)), 156)
    err = (out - ref_out).abs().max().item() / ref_out.abs().max().item()
    show(f"scatter binned accum=0, {mb} MiB bricks",
         t(lambda: call("ndjir_voxel_grad_feature_binned", B, gf, go, q, [G] * 3, D, MN, MX, 0, ws, wsb, 0)), 284)
    err2 = (gf - ref_gf).abs().max().item() / ref_gf.abs().max().item()
    show(f"scatter binned accum=1, {mb} MiB bricks",
         t(lambda: call("ndjir_voxel_grad_feature_binned", B, gf, go, q, [G] * 3, D, MN, MX, 1, ws, wsb, 0)), 284)
    print(f"    max-norm rel. diff vs direct: fwd {err:.2e}, grad_feature {err2:.2e}", flush=True)
call("ndjir_set_option", "voxel_bin_mb", 16)
call("ndjir_set_option", "voxel_tma", 0)
show("gather binned, 16 MiB bricks, L2-window sweep (voxel_tma=0)",
     t(lambda: call("ndjir_voxel_query_on_voxel_binned", B, out, q, feat, [G] * 3, D, MN, MX, 0, ws, wsb, 0)), 156)
print(f"    max-norm rel. diff vs direct: {(out - ref_out).abs().max().item() / ref_out.abs().max().item():.2e}")
call("ndjir_set_option", "voxel_tma", 1)
call("ndjir_set_option", "voxel_pair256", 1)
show("gather binned, 16 MiB, 256-bit z-pair loads",
     t(lambda: call("ndjir_voxel_query_on_voxel_binned", B, out, q, feat, [G] * 3, D, MN, MX, 0, ws, wsb, 0)), 156)
call("ndjir_set_option", "voxel_pair256", 0)
call("ndjir_set_option", "voxel_binned", -1)
show("gather auto (cudaMallocAsync scratch)", t(lambda: call("ndjir_voxel_query_on_voxel", B, out, q, feat, [G] * 3, D, MN, MX, 0, 0)), 156)
show("scatter auto accum=0", t(lambda: call("ndjir_voxel_grad_feature", B, gf, go, q, [G] * 3, D, MN, MX, 0, 0)), 284)
