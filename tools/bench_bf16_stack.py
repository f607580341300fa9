"""BASELINE.json configs[2]: the per-point MLP stack at B=64 x N=8192 in bf16 (P2C_PREC_BF16) next to 3xTF32 -
every tensor-core layer shape of the backbone at that batch, timed back to back (L2 flushed), for the ncu
tensor-pipe capture:  ncu --set full -k regex:linear_tc_ss python tools/bench_bf16_stack.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops

B, N = 64, 8192
LAYERS = [("sa1.1", B*512*64, 64, 64, 0), ("sa1.2", B*512*64, 64, 128, 64), ("sa2.1", B*128*64, 128, 128, 0),
          ("sa2.2", B*128*64, 128, 256, 64), ("sa3.0", B*128, 259, 256, 0), ("sa3.1", B*128, 256, 512, 0),
          ("sa3.2", B*128, 512, 1024, 128), ("fp3.0", B*128, 1280, 256, 0), ("fp3.1", B*128, 256, 256, 0),
          ("fp2.0", B*512, 384, 256, 0), ("fp2.1", B*512, 256, 128, 0), ("fp1.x", B*N, 128, 128, 0)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
tot = {1: 0.0, 2: 0.0}
flops = 0.0
for name, M, K, Nn, pool in LAYERS:
    ld = ops.pad4(K)
    X = torch.randn(M, ld, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
    Y = torch.empty(M, Nn, device="cuda") if pool == 0 else None
    line = f"{name:6s} M={M:8d} K={K:4d} N={Nn:4d}"
    for prec in (_lib.PREC_3XTF32, _lib.PREC_BF16):
        wop = ops.weight_operand(X, W, Nn, K, False, pool, prec)
        ts = []
        for it in range(3):
            flush.zero_(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_(); s.record()
            for _ in range(4):
                ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, out=Y, want_y=(pool == 0),
                           precision=prec, w_split=wop)
            e.record(); e.synchronize(); ts.append(s.elapsed_time(e) / 4)
        t = min(ts[1:]); tot[prec] += t
        line += f" | {'3xtf32' if prec == 1 else 'bf16  '} {t*1e3:7.1f} us {2.0*M*K*Nn/t/1e9:7.1f} TF/s {4.0*M*(K+(0 if pool else Nn))/t/1e6:6.0f} GB/s"
    flops += 2.0 * M * K * Nn
    print(line, flush=True)
print(f"stack total: 3xtf32 {tot[1]:.3f} ms ({flops/tot[1]/1e9:.0f} TFLOP/s)   bf16 {tot[2]:.3f} ms ({flops/tot[2]/1e9:.0f} TFLOP/s)")
