import numpy as np, torch, sys
sys.path.insert(0, "/root/repo")
from ndjir_b200 import _lib
def r4(n): return (n + 3) // 4 * 4
def dev(a): return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
for (rows, K_in, N, split) in [(2048, 259, 213, 1), (2048, 256, 213, 1), (2048, 259, 224, 1), (2048, 259, 256, 1)]:
  for fused in (1, 0):
    for nanpad in (1, 0):
        rng = np.random.RandomState(rows + N)
        A = rng.randn(rows, K_in).astype(np.float32); dZ = rng.randn(rows, N).astype(np.float32)
        lda, ldz, ldw = r4(K_in) + 4, r4(N) + 4, r4(N) + 8
        fillv = np.nan if nanpad else 0.0
        As = np.full((rows, lda), fillv, np.float32); As[:, :K_in] = A
        Zs = np.full((rows, ldz), fillv, np.float32); Zs[:, :N] = dZ
        gW0 = np.zeros((K_in, ldw), np.float32); gb0 = np.zeros(r4(N) + 4, np.float32)
        dA, dZd, gW, gb = dev(As), dev(Zs), dev(gW0), dev(gb0)
        _lib.call("ndjir_set_option", "mlp_fused_colsum", fused)
        _lib.call("ndjir_wgrad_bias", rows, K_in, N, dA, lda, dZd, ldz, gW, ldw, gb, split, 0)
        torch.cuda.synchronize()
        wantW = A.astype(np.float64).T @ dZ.astype(np.float64); wantb = dZ.astype(np.float64).sum(0)
        gotW, gotb = gW.cpu().numpy(), gb.cpu().numpy()
        eW = np.abs(gotW[:, :N] - wantW); eb = np.abs(gotb[:N] - wantb)
        print(rows, K_in, N, "fused", fused, "nanpad", nanpad, "nanW", int(np.isnan(gotW).sum()), "errW", np.nanmax(eW) / np.abs(wantW).max(), "errb", np.nanmax(eb) / np.abs(wantb).max(), "bad rows", np.unique(np.where(~(eW < 1e-3))[0])[:6], "bad cols", np.unique(np.where(~(eW < 1e-3))[1])[:6])
