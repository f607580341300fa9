import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops
M, K, Nn = 262144, 128, 128
ld = ops.pad4(K)
X = torch.randn(M, ld, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
ts = []
for it in range(6):
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=0, want_y=True, precision=1)
    e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
print("P2C_TC_DBG", os.environ.get("P2C_TC_DBG"), "%.1f us" % (min(ts[1:]) * 1e3))
