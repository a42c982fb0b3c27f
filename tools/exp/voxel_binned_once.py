"""One binned gather + one binned scatter at the micro-benchmark shape (for ncu launch lists)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndjir_b200._lib import call  # noqa: E402
B, G, D = 1 << 24, 512, 4
MN, MX = [-1.0] * 3, [1.0] * 3
q = torch.rand(B, 3, device="cuda") * 2 - 1
feat = torch.randn(G, G, G, D, device="cuda") * 0.01
out = torch.empty(B, D, device="cuda")
go = torch.ones(B, D, device="cuda")
gf = torch.zeros_like(feat)
wsb = call("ndjir_voxel_binned_workspace_bytes", B)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
for _ in range(2):
    call("ndjir_voxel_query_on_voxel_binned", B, out, q, feat, [G] * 3, D, MN, MX, 0, ws, wsb, 0)
    call("ndjir_voxel_grad_feature_binned", B, gf, go, q, [G] * 3, D, MN, MX, 1, ws, wsb, 0)
    call("ndjir_set_option", "voxel_binned", 0)
    call("ndjir_voxel_query_on_voxel", B, out, q, feat, [G] * 3, D, MN, MX, 0, 0)
    call("ndjir_voxel_grad_feature", B, gf, go, q, [G] * 3, D, MN, MX, 1, 0)
    call("ndjir_set_option", "voxel_binned", -1)
torch.cuda.synchronize()
