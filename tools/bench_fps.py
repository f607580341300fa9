import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import ops, synthetic
B, N = 32, 8192
xyz = synthetic.s_cyl(B, N, 8, 1234)["pcs"].cuda()
start = torch.zeros(B, dtype=torch.long, device="cuda")
for it in range(3):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        idx, nx = ops.fps(xyz, 512, start)
    e.record(); e.synchronize()
print("P2C_FPS_PPT", os.environ.get("P2C_FPS_PPT"), "fps L1 %.1f us" % (s.elapsed_time(e) / 5 * 1e3), int(idx.sum()))
