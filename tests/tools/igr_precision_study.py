"""CPU study for the next block: which tensor-core operand format can carry the implicit
network (8 x 512, softplus beta=100) and its input gradient within the 1e-4 bar?  Emulates the operand roundings of
a GEMM y = x W^T with fp32 accumulation:
  fp32      reference arithmetic of the framework (what the reference computes on CPU)
  3xtf32    x = x_hi + x_lo, W = W_hi + W_lo (hi = low 13 mantissa bits cleared, lo = remainder rounded to tf32),
            y = x_hi W_hi + x_lo W_hi + x_hi W_lo          (the scheme of csrc/linear_tc.cu)
  tf32      y = x_hi W_hi
  bf16      operands rounded to bfloat16
against a float64 evaluation of the same weights.  Prints max relative errors of f and of df/dx[:, -2:]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import math  # noqa: E402

import torch  # noqa: E402

from oracle import igr_oracle as igr  # noqa: E402


def tf32_hi(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def gemm(x, W, mode):
    if mode == "fp32":
        return x @ W.t()
    if mode == "bf16":
        return (x.bfloat16().float()) @ (W.bfloat16().float()).t()
    xh, Wh = tf32_hi(x), tf32_hi(W)
    if mode == "tf32":
        return xh @ Wh.t()
    xl, Wl = tf32_hi(x - xh), tf32_hi(W - Wh)
    return xh @ Wh.t() + xl @ Wh.t() + xh @ Wl.t()


def run(sd, x, mode, beta=100.0, skip=(4,)):
    L = len([k for k in sd if k.endswith(".weight")]) - 1
    d_in = x.shape[1]
    h, zs = x, []
    for i in range(L + 1):
        p = torch.cat([h, x], -1) / math.sqrt(2.0) if i in skip else h
        z = gemm(p, sd[f"lin{i}.weight"], mode) + sd[f"lin{i}.bias"]
        zs.append(z)
        h = igr.softplus(z, beta) if i < L else z
    g = torch.ones_like(zs[-1])
    gx = torch.zeros_like(x)
    for i in range(L, -1, -1):
        if i < L:
            bz = beta * zs[i]
            g = g * torch.where(bz > 20, torch.ones_like(bz), torch.sigmoid(bz))
        g = gemm(g, sd[f"lin{i}.weight"].t().contiguous(), mode)
        if i in skip:
            g = g / math.sqrt(2.0)
            gx = gx + g[:, -d_in:]
            g = g[:, :-d_in]
    return zs[-1], (gx + g)[:, -2:]


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    sd = igr.implicit_init(seed=0)
    R = 4096
    lat = torch.nn.functional.normalize(torch.randn(R // 64, 256), dim=-1).repeat_interleave(64, 0)
    pts = torch.rand(R, 2) * 2 - 1
    x = torch.cat([lat, pts], -1)
    sd64 = {k: v.double() for k, v in sd.items()}
    f64, g64 = igr.implicit_forward_with_input_grad(sd64, x.double())
    print(f"rows {R}; |f| max {float(f64.abs().max()):.3f}, |grad| max {float(g64.abs().max()):.3f}")
    for mode in ("fp32", "3xtf32", "tf32", "bf16"):
        f, g = run(sd, x, mode)
        ef = float((f.double() - f64).abs().max() / f64.abs().max())
        eg = float((g.double() - g64).abs().max() / g64.abs().max())
        print(f"{mode:7s} rel err  f {ef:.2e}   df/dx {eg:.2e}")
