"""Per-kernel times of one brick-ordered voxel gather at the micro-benchmark shape (torch.profiler)."""
import os, sys, collections
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndjir_b200._lib import call  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
B, G, D = 1 << 24, 512, 4
MN, MX = [-1.0] * 3, [1.0] * 3
rng = np.random.RandomState(412)
q = torch.as_tensor((rng.rand(B, 3) * 2 - 1).astype(np.float32)).cuda()
feat = torch.randn(G, G, G, D, device="cuda") * 0.01
out = torch.empty(B, D, device="cuda")
wsb = call("ndjir_voxel_binned_workspace_bytes", B)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tma, bx, l2, dbg in ((1, 16, 1, 0), (1, 8, 1, 0), (0, 16, 1, 0)):
    call("ndjir_set_option", "voxel_tma_dbg", dbg)
    call("ndjir_set_option", "voxel_tma", tma)
    call("ndjir_set_option", "voxel_tma_bx", bx)
    call("ndjir_set_option", "voxel_tma_l2", l2)
    for _ in range(2):
        call("ndjir_voxel_query_on_voxel_binned", B, out, q, feat, [G] * 3, D, MN, MX, 0, ws, wsb, 0)
    flush.zero_()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        call("ndjir_voxel_query_on_voxel_binned", B, out, q, feat, [G] * 3, D, MN, MX, 0, ws, wsb, 0)
        torch.cuda.synchronize()
    print(f"voxel_tma={tma} bx={bx} l2={l2} dbg={dbg}")
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            if ev.device_time_total > 100:
                print(f"   {ev.device_time_total / 1e3:8.3f} ms  {ev.name[:60]}")
