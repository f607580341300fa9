import sys, torch
sys.path.insert(0, "/root/repo")
from point2cyl_b200 import ops, synthetic, _lib
B, N = 128, 32768
xyz = synthetic.s_uniform(B, N, seed=77).cuda()
s1 = torch.zeros(B, dtype=torch.long, device="cuda")
def step():
    _, c1 = ops.fps(xyz, 512, s1); g1 = ops.ball_query(0.2, 64, xyz, c1); _, c2 = ops.fps(c1, 128, s1); g2 = ops.ball_query(0.4, 64, c1, c2)
for _ in range(3): step()
torch.cuda.synchronize()
_lib.profile_start(); step(); print([(n, round(t, 3)) for n, tag, t in _lib.profile_stop()])
