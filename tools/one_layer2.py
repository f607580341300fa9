"""One tcgen05 layer launch at a given shape (for ncu): one_layer2.py M K N pool"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import ops
M, K, Nn, pool = (int(v) for v in sys.argv[1:5])
X = torch.randn(M, K, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
for _ in range(3):
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
    ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, want_y=(pool == 0), precision=1)
torch.cuda.synchronize()
