"""Time the big tcgen05 layers with parts of the kernel disabled (P2C_TC_DBG bit0: no transform math, bit1: no
epilogue body) to see which role paces the pipeline."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for mode in ("0", "1", "2", "3"):
        env = dict(os.environ, P2C_TC_DBG=mode)
        print("P2C_TC_DBG =", mode, flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=env)
    sys.exit(0)
import torch
from point2cyl_b200 import _lib, ops
B, N = 32, 8192
LAYERS = [("sa1.1", B*512*64, 64, 64, 0), ("sa1.2", B*512*64, 64, 128, 64), ("sa2.1", B*128*64, 128, 128, 0),
          ("sa2.2", B*128*64, 128, 256, 64), ("fp1.x", B*N, 128, 128, 0)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, M, K, Nn, pool in LAYERS:
    X = torch.randn(M, K, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
    Y = torch.empty(M, Nn, device="cuda") if pool == 0 else None
    wsp = ops.split_tf32(W) if os.environ.get("P2C_TC_FORCE_SS") else None
    ts = []
    for it in range(3):
        flush.zero_(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); s.record()
        for _ in range(8):
            ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, out=Y, want_y=(pool == 0), precision=1,
                       w_split=wsp)
        e.record(); e.synchronize(); ts.append(s.elapsed_time(e) / 8)
    print(f"  {name} M={M} K={K} N={Nn} pool={pool}: {min(ts[1:])*1e3:.1f} us", flush=True)
