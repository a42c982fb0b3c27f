"""Experiment: cost of the voxel-hash forward / grad_feature per level (bench shape: G0=16, gf=1.5, T0=2^15, D=2)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndjir_b200._lib import call
B, D, T0 = 1 << 24, 2, 2 ** 15
MN, MX = [-1.0] * 3, [1.0] * 3
q = torch.rand(B, 3, device="cuda") * 2 - 1
def t(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for lvl in (0, 1, 2, 3, 4, 6, 8, 12, 15):
    G = int(np.floor(16 * 1.5 ** lvl))
    n = call("ndjir_voxel_hash_num_params", G, 1.0, T0, 1, D)
    feat = torch.randn(n, device="cuda") * 0.01
    out = torch.empty(D * B, device="cuda"); go = torch.ones(D * B, device="cuda"); gf = torch.zeros_like(feat)
    f = t(lambda: call("ndjir_voxel_hash_voxel_hash_feature", B, out, q, feat, G, 1.0, T0, 1, D, MN, MX, 0, 0, 0))
    b = t(lambda: call("ndjir_voxel_hash_grad_feature", B, gf, go, q, G, 1.0, T0, 1, D, MN, MX, 0, 1, 0))
    print(f"level {lvl:2d} G={G:5d} T={min(G**3, T0):6d}: fwd {f:.3f} ms  grad_feature {b:.3f} ms", flush=True)


# full bench shape: grad_feature with / without the shared-memory privatised coarse levels
G0, gf_, L = 16, 1.5, 16
n = call("ndjir_voxel_hash_num_params", G0, gf_, T0, L, D)
gfe = torch.zeros(n, device="cuda"); go = torch.ones(D * L * B, device="cuda")
for mode in (0, 1, 0, 1):
    call("ndjir_set_option", "hash_coarse_private", mode)
    ms = t(lambda: call("ndjir_voxel_hash_grad_feature", B, gfe, go, q, G0, gf_, T0, L, D, MN, MX, 0, 0, 0))
    print(f"bench shape grad_feature, hash_coarse_private={mode}: {ms:.3f} ms", flush=True)
call("ndjir_set_option", "hash_coarse_private", 1)
