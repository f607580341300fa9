"""Two passes over the tensor-core layers of the MLP stack at B=64 x N=8192 in bf16 (for ncu: skip the first 12)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops
B, N = 64, 8192
LAYERS = [("sa1.1", B*512*64, 64, 64, 0), ("sa1.2", B*512*64, 64, 128, 64), ("sa2.1", B*128*64, 128, 128, 0),
          ("sa2.2", B*128*64, 128, 256, 64), ("sa3.0", B*128, 259, 256, 0), ("sa3.1", B*128, 256, 512, 0),
          ("sa3.2", B*128, 512, 1024, 128), ("fp3.0", B*128, 1280, 256, 0), ("fp3.1", B*128, 256, 256, 0),
          ("fp2.0", B*512, 384, 256, 0), ("fp2.1", B*512, 256, 128, 0), ("fp1.x", B*N, 128, 128, 0)]
data = []
for name, M, K, Nn, pool in LAYERS:
    X = torch.randn(M, ops.pad4(K), device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
    data.append((X, W, torch.randn(Nn, device="cuda"), torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda"),
                 ops.weight_operand(X, W, Nn, K, False, pool, _lib.PREC_BF16), K, Nn, pool))
for _ in range(2):
    for X, W, b, sc, sh, wop, K, Nn, pool in data:
        stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda")
        ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, want_y=(pool == 0),
                   precision=_lib.PREC_BF16, w_split=wop)
torch.cuda.synchronize()
