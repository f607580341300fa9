import os, sys
sys.path.insert(0, "/root/repo")
import torch
from point2cyl_b200 import _lib, ops
B, N = 32, 8192
MODES = sys.argv[1:] or ["0", "64", "128"]
LAYERS = [("sa1.1", B*512*64, 64, 64, 0), ("sa1.2", B*512*64, 64, 128, 64), ("sa2.1", B*128*64, 128, 128, 0),
          ("sa2.2", B*128*64, 128, 256, 64), ("fp1.x", B*N, 128, 128, 0)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, M, K, Nn, pool in LAYERS:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
    Y = torch.empty(M, Nn, device="cuda") if pool == 0 else None
    res = []
    for mode in MODES:
        os.environ["P2C_TC_DBG"] = mode
        ts = []
        for it in range(4):
            flush.zero_(); torch.cuda._sleep(2000000)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, out=Y, want_y=(pool == 0), precision=1)
            e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
        res.append(min(ts[1:]) * 1e3)
    print(f"{name} M={M} K={K} N={Nn} pool={pool}: " + "  ".join(f"dbg {m}: {r:.1f} us" for m, r in zip(MODES, res)), flush=True)
