"""A few p2c_sa_xyz_linear launches at the sa1 shape of config 2 (for ncu captures): one_sa_xyz.py [N1 pool]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import ops, synthetic
N1, pool = (int(v) for v in (sys.argv[1:3] or (64, 0)))
B, N, S, ns = 32, 8192, 512, 64
xyz = synthetic.s_cyl(B, N, 8, 1)["pcs"].cuda()
_, cxyz = ops.fps(xyz, S, torch.zeros(B, dtype=torch.long, device="cuda"))
gidx = ops.ball_query(0.2, ns, xyz, cxyz)
g = torch.Generator().manual_seed(0)
W0, b0 = torch.randn(64, 3, generator=g).cuda(), torch.randn(64, generator=g).cuda()
sc, sh = (torch.rand(64, generator=g) + 0.5).cuda(), torch.randn(64, generator=g).cuda()
W1, b1 = (torch.randn(N1, 64, generator=g) / 8).cuda(), torch.randn(N1, generator=g).cuda()
for _ in range(3):
    st = torch.zeros(2 * N1, dtype=torch.float64, device="cuda")
    ops.sa_xyz_linear(xyz, cxyz, gidx, W0, b0, W1, b1, scale0=sc, shift0=sh, stats=st, pool_group=pool, want_y=(pool == 0))
torch.cuda.synchronize()
print("done")
