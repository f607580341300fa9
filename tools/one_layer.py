"""Run one p2c_linear shape a few times (for ncu captures): one_layer.py M K N pool prec [want_y]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops
M, K, Nn, pool, prec = (int(v) for v in sys.argv[1:6])
ld = ops.pad4(K)
X = torch.randn(M, ld, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
for it in range(3):
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
    ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, want_y=(pool == 0), precision=prec)
torch.cuda.synchronize()
print("done")
