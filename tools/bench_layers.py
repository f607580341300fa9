"""Per-layer timing of p2c_linear at BASELINE config 2 shapes, both precisions (GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops

B, N = 32, 8192
LAYERS = [("sa1.0", B*512*64, 3, 64, 0), ("sa1.1", B*512*64, 64, 64, 0), ("sa1.2", B*512*64, 64, 128, 64),
          ("sa2.0", B*128*64, 131, 128, 0), ("sa2.1", B*128*64, 128, 128, 0), ("sa2.2", B*128*64, 128, 256, 64),
          ("sa3.0", B*128, 259, 256, 0), ("sa3.1", B*128, 256, 512, 0), ("sa3.2", B*128, 512, 1024, 128),
          ("fp3.0", B*128, 1280, 256, 0), ("fp3.1", B*128, 256, 256, 0), ("fp2.0", B*512, 384, 256, 0),
          ("fp2.1", B*512, 256, 128, 0), ("fp1.x", B*N, 128, 128, 0), ("fc2", B*N, 128, 19, 0)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, M, K, Nn, pool in LAYERS:
    ld = ops.pad4(K)
    X = torch.randn(M, ld, device="cuda")
    W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    b = torch.randn(Nn, device="cuda")
    sc = torch.rand(K, device="cuda") + 0.5
    sh = torch.randn(K, device="cuda")
    out = []
    for prec in (_lib.PREC_FP32, _lib.PREC_3XTF32):
        path = _lib.load().p2c_linear_path(ld, M, Nn, K, 0, pool, prec, 1)
        wsp = ops.split_tf32(W) if path == 2 else None
        ts = []
        REP = 8
        stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
        Y = torch.empty(M, Nn, device="cuda") if pool == 0 else None
        for it in range(3):
            flush.zero_()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # a long-running kernel first so the launches below queue up behind it (hides host latency)
            flush.zero_()
            s.record()
            for _ in range(REP):
                ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, out=Y, want_y=(pool == 0), precision=prec, w_split=wsp)
            e.record(); e.synchronize()
            ts.append(s.elapsed_time(e) / REP)
        t = min(ts[1:])
        fl = 2.0 * M * K * Nn
        by = 4.0 * M * (K + (0 if pool else Nn))
        out.append(f"path{path} {t*1e3:8.1f} us {fl/t/1e9:7.1f} TF/s {by/t/1e6:7.0f} GB/s")
    print(f"{name:6s} M={M:8d} K={K:4d} N={Nn:4d} | " + " | ".join(out), flush=True)
