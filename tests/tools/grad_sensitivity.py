"""How well-conditioned are the reference's parameter gradients?  (CPU, oracle autograd.)

Perturbs every weight by a relative 1e-6 and reports how far the gradients of one training step move, with
train-mode BatchNorm and with BatchNorm on running statistics.  This is the bound on how tightly ANY implementation
with a different summation order can reproduce the reference's gradients (tests/test_gpu_backward.py cites it).

    python tests/tools/grad_sensitivity.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import p2c_oracle as orc  # noqa: E402
from point2cyl_b200 import synthetic  # noqa: E402


def grads(B, N, K, seed, training, noise):
    data = synthetic.s_cyl(B, N, K, seed)
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0
    torch.manual_seed(seed)
    starts = (torch.randint(0, N, (B,)), torch.randint(0, 512, (B,)))
    gen = torch.Generator().manual_seed(77)
    sd = {k: ((v * (1 + noise * torch.randn(v.shape, generator=gen))).clone().requires_grad_(True)
              if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in orc.init_state_dict((3, 2 * K), seed=seed).items()}
    orc.forward_loss(sd, data, training=training, fps_start=starts, dropout_mask=mask)["total"].backward()
    return {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}


def forward_outputs(B, N, K, seed, training, noise):
    data = synthetic.s_cyl(B, N, K, seed)
    torch.manual_seed(seed)
    starts = (torch.randint(0, N, (B,)), torch.randint(0, 512, (B,)))
    gen = torch.Generator().manual_seed(5)
    sd = {k: (v * (1 + noise * torch.randn(v.shape, generator=gen)) if v.is_floating_point() and "running" not in k
              else v.clone()) for k, v in orc.init_state_dict((3, 2 * K), seed=seed).items()}
    with torch.no_grad():
        return orc.backbone_forward(sd, data["pcs"], training=training, fps_start=starts,
                                    dropout_mask=torch.ones(B, 128, N))


if __name__ == "__main__":
    for B, training in ((2, True), (1, True), (2, False)):
        (X0, W0), (X1, W1) = forward_outputs(B, 1024, 4, 0, training, 0.0), forward_outputs(B, 1024, 4, 0, training, 1e-6)
        print(f"FORWARD B={B} N=1024 train-mode BN={training}: output change under a 1e-6 relative weight perturbation: "
              f"X {float((X1 - X0).abs().max() / X0.abs().max()):.1e}  W_raw {float((W1 - W0).abs().max() / W0.abs().max()):.1e}")
    KEYS = ("fc2.1.weight", "bn1.bias", "fc1.weight", "fp1.mlp_convs.0.weight", "fp2.mlp_convs.0.weight",
            "sa2.mlp_convs.1.weight", "sa1.mlp_convs.1.weight", "sa1.mlp_bns.0.bias")
    for B, training in ((2, True), (8, True), (2, False)):
        a, b = grads(B, 1024, 4, 0, training, 0.0), grads(B, 1024, 4, 0, training, 1e-6)
        print(f"B={B} N=1024 K=4 train-mode BN={training}: gradient change under a 1e-6 relative weight perturbation")
        for k in KEYS:
            d = ((a[k] - b[k]).abs() / a[k].abs().max()).reshape(-1)
            q = torch.quantile(d, torch.tensor([0.5, 0.99, 1.0]))
            print(f"  {k:26s} median {float(q[0]):.1e}  p99 {float(q[1]):.1e}  max {float(q[2]):.1e}  "
                  f"L2 {float((a[k] - b[k]).norm() / a[k].norm()):.1e}")
