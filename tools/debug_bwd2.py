"""Isolate gradient error: d_out of the loss kernels vs oracle autograd on the kernel's own network outputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import p2c_oracle as orc
from point2cyl_b200 import pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone

g = np.load("tests/golden/train_b2_n1024_k4.npz")
B, N, K, seed = (int(v) for v in g["meta"])
DEV = "cuda"
cpu = synthetic.s_cyl(B, N, K, seed)
data = {k: v.to(DEV) for k, v in cpu.items()}
Xr = torch.from_numpy(g["X_raw"]); Wr = torch.from_numpy(g["W_raw"])
for w in [(1, 1, 1, 1, 1), (1, 0, 0, 0, 0), (0, 1, 0, 0, 0), (0, 0, 1, 0, 0), (0, 0, 0, 1, 0), (0, 0, 0, 0, 1)]:
    Xc, Wc = Xr.clone().requires_grad_(True), Wr.clone().requires_grad_(True)
    ref = orc.loss_block(cpu["pcs"], Xc, Wc, cpu["normals"], cpu["inst"], cpu["bb"], cpu["axes"], cpu["centers"], weights=w)
    ref["total"].backward()
    Xd, Wd = Xr.to(DEV).requires_grad_(True), Wr.to(DEV).requires_grad_(True)
    out = pipeline.loss_forward(data["pcs"], Xd, Wd, data["normals"], data["inst"], data["bb"], data["axes"], data["centers"], weights=w)
    out["total"].backward()
    for name, got, want in (("dX", Xd.grad.cpu(), Xc.grad), ("dW", Wd.grad.cpu(), Wc.grad)):
        d = (got - want).double()
        print(w, name, "max rel %.2e" % float(d.abs().max() / want.abs().max().clamp_min(1e-30)),
              "l2 rel %.2e" % float(d.norm() / want.double().norm().clamp_min(1e-30)),
              "colsum rel %.2e" % float(d.sum((0, 1)).abs().max() / want.double().sum((0, 1)).abs().max().clamp_min(1e-30)))
    print("   match equal", torch.equal(out["matching_indices"].cpu(), ref["matching_indices"]), "E_AX", out["E_AX"][0, 0].cpu().numpy(), ref["E_AX"][0, 0].detach().numpy())
