"""Per-parameter gradient error of one training step against the reference golden (debug helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import p2c_oracle as orc
from point2cyl_b200 import pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone

g = np.load("tests/golden/train_b2_n1024_k4.npz")
B, N, K, seed = (int(v) for v in g["meta"])
DEV = "cuda"
data = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed).items()}
net = backbone(output_sizes=[3, 2 * K])
net.load_state_dict(orc.init_state_dict((3, 2 * K), seed=seed), strict=True)
net = net.to(DEV).train()
mask = ((torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0).to(DEV)
pipeline.dropout_mask_fn = lambda x, p=0.5, **kw: mask
if len(sys.argv) > 1:
    pipeline.set_precision(sys.argv[1])
starts = (torch.from_numpy(g["s1"]).to(DEV), torch.from_numpy(g["s2"]).to(DEV))
X_raw, W_raw = net(data["pcs"], fps_start=starts)
out = pipeline.loss_forward(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"], data["axes"], data["centers"])
out["total"].backward()
for k, p in reversed(list(net.named_parameters())):
    ref = torch.from_numpy(g["grad_" + k]).double()
    got = p.grad.reshape(-1).double().cpu()
    err = float((got[:ref.numel()] - ref).abs().max())
    print(f"{k:28s} ref_max {float(ref.abs().max()):.3e} err {err:.3e} rel {err / max(float(ref.abs().max()), 1e-30):.2e} "
          f"norm {float(got.norm()):.4e} ref_norm {float(g['gnorm_' + k]):.4e}")
