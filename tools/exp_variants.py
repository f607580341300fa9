import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops
def run(M, K, Nn, pool, want_y, use_stats, affine, prec=1):
    ld = ops.pad4(K)
    X = torch.randn(M, ld, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    ts = []
    for it in range(5):
        stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda") if use_stats else None
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.linear(X, W, b, K=K, in_scale=sc if affine else None, in_shift=sh if affine else None, stats=stats,
                   pool_group=pool, want_y=want_y, precision=prec)
        e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts[1:]) * 1e3
for (M, K, Nn) in [(262144, 128, 128), (1048576, 64, 64), (262144, 128, 64), (262144, 64, 128), (262144, 32, 128)]:
    print(M, K, Nn, "Y+stats %.0f | Y only %.0f | stats only %.0f | pool64+stats %.0f | Y only no-affine %.0f us" % (
        run(M, K, Nn, 0, True, True, True), run(M, K, Nn, 0, True, False, True), run(M, K, Nn, 0, False, True, True),
        run(M, K, Nn, 64, False, True, True), run(M, K, Nn, 0, True, False, False)), flush=True)
