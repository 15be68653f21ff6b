import numpy as np, torch, sys, os
sys.path.insert(0, '.')
from graphaibench_b200 import ops
from graphaibench_b200._abi import lib
np.set_printoptions(linewidth=200, precision=3, suppress=True)
def run(X, G, mode=3, out=None):
    lib().gai_set_gemm_mode(mode)
    o = ops.matmul(torch.from_numpy(X).cuda(), torch.from_numpy(G).cuda(), transA=True, out=out, accum=False).cpu().numpy()
    lib().gai_set_gemm_mode(0)
    return o
n, kx, my = 4096, 128, 64
# dirty TMEM with the forward kernel first
A = torch.randn(8192, 128, device='cuda'); W = torch.randn(128, 64, device='cuda')
lib().gai_set_gemm_mode(2); f = ops.matmul(A, W); lib().gai_set_gemm_mode(0)
print("fwd err", float((f - A @ W).abs().max()))
X = np.ones((n, kx), np.float32); G = np.ones((n, my), np.float32)
out = torch.full((kx, my), 7.0, device='cuda')
o = run(X, G, 3, out); print(os.environ.get("GAI_DBG_IDESC"), os.environ.get("GAI_DBG_LBO"), os.environ.get("GAI_DBG_SBO"), "ones: nnz", np.count_nonzero(o), "uniq", np.unique(o)[:10])
X = np.tile(np.arange(1, kx + 1, dtype=np.float32), (n, 1))
o = run(X, G, 3); print("row pattern:", o[:8, 0] / n)
X = np.ones((n, kx), np.float32); G = np.tile(np.arange(1, my + 1, dtype=np.float32), (n, 1))
o = run(X, G, 3); print("col pattern:", o[0, :8] / n)
