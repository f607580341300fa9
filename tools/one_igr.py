"""A few implicit-network layer launches for ncu (forward sweep + input-gradient sweep [+ backward] at I instances x
(1024 + 1152) points):  ncu --set full -k regex:linear_tc_ss -s 8 -c 4 python tools/one_igr.py 64"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point2cyl_b200 import igr  # noqa: E402
from point2cyl_b200.dropin.IGR import network as dnet  # noqa: E402

I = int(sys.argv[1]) if len(sys.argv) > 1 else 64
backward = len(sys.argv) > 2 and sys.argv[2] == "bwd"
S = 1024
dev = torch.device("cuda")
torch.manual_seed(0)
net = dnet.ImplicitNet(d_in=258, dims=[512] * 8, skip_in=[4], geometric_init=True, radius_init=1, beta=100).to(dev)
latent = torch.nn.functional.normalize(torch.randn(I, 256, device=dev), dim=1)
on = torch.rand(I, S, 2, device=dev) * 2 - 1
off = torch.rand(I, S + S // 8, 2, device=dev) * 3.6 - 1.8
for _ in range(2):
    f, ctx = igr.implicit_forward(net, latent=latent, pts=[on, off])
    g = igr.implicit_input_gradient(ctx)
    if backward:
        igr.implicit_backward(ctx, torch.randn_like(f) / ctx.R, torch.randn_like(g) / ctx.R)
torch.cuda.synchronize()
print("ok", float(f.abs().mean()), float(g.abs().mean()))
