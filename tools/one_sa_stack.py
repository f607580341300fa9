"""A few launches of the one-kernel SA level at config-2 size for ncu:
ncu --set full -k regex:sa_stack -c 1 python tools/one_sa_stack.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point2cyl_b200 import pipeline, synthetic  # noqa: E402
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone  # noqa: E402

B, N, K = 32, 8192, 8
dev = torch.device("cuda")
data = synthetic.s_cyl(B, N, K, seed=1234)
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to(dev).eval()
pcs = data["pcs"].to(dev)
start = [torch.randint(0, N, (B,)).to(dev), torch.randint(0, 512, (B,)).to(dev)]
geo = pipeline.geometry_forward(net, pcs, start, moments=False)
with torch.no_grad():
    for _ in range(3):
        _, out = pipeline.set_abstraction(net.sa1, geo.xyz, None, None, geo=(geo.fps1, geo.l1_xyz, geo.gidx1))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
