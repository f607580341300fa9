"""GPU box: per-shape errors of p2c_linear_act and per-layer errors of the implicit network against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import math
import torch
import torch.nn.functional as F
from oracle import igr_oracle as orc
from point2cyl_b200 import igr, ops
from point2cyl_b200.dropin.IGR import network as dnet

DEV = "cuda"


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


g = torch.Generator().manual_seed(0)
for M, N, K in [(1000, 512, 512), (300, 254, 512), (513, 512, 254), (2048, 512, 258), (128, 128, 64)]:
    ld = ops.pad4(K)
    X = torch.zeros(M, ld); X[:, :K] = torch.randn(M, K, generator=g) * 0.05
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.01
    Xd, Wd = X.to(DEV), W.to(DEV)
    ws = ops.split_tf32_multi([Wd])[0]
    Z = X[:, :K].double() @ W.double().t() + b.double()
    out = torch.zeros(M, ops.pad4(N), device=DEV); S = torch.zeros(M, ops.pad4(N), device=DEV)
    ops.linear_act(Xd, ws, b.to(DEV), N, K, op=1, beta=100.0, oscale=0.5, out=out[:, :N], S=S[:, :N])
    e1 = rel(out[:, :N], 0.5 * F.softplus(Z, beta=100.0))
    e2 = rel(S[:, :N], torch.sigmoid(100 * Z))
    o0 = ops.linear_act(Xd, ws, b.to(DEV), N, K, op=0)
    e0 = rel(o0, Z)
    print(f"M={M} N={N} K={K}: plain {e0:.2e} softplus {e1:.2e} sigmoid {e2:.2e}", flush=True)

net = dnet.ImplicitNet(d_in=258, dims=[512] * 8, skip_in=[4]).to(DEV)
sd = orc.implicit_init(seed=5)
net.load_state_dict(sd)
I, S_ = 3, 64
latent = F.normalize(torch.randn(I, 256, generator=g), dim=1)
on = torch.rand(I, S_, 2, generator=g) * 2 - 1
x = orc.add_latent(on, latent)
# oracle activations
h, zs, hs = x, [], []
for i in range(9):
    if i == 4:
        h = torch.cat([h, x], -1) / math.sqrt(2)
    z = F.linear(h, sd[f"lin{i}.weight"], sd[f"lin{i}.bias"])
    zs.append(z)
    h = orc.softplus(z) if i < 8 else z
    hs.append(h)
# kernel layer by layer with the oracle's input of each layer (isolates the layer)
lins = igr._layers(net)
wsplit = ops.split_tf32_multi([l.weight for l in lins[:8]])
hin = x
for i in range(8):
    if i == 4:
        hin = torch.cat([hs[3], x], -1) / math.sqrt(2)
    elif i > 0:
        hin = hs[i - 1]
    out_i, in_i = lins[i].weight.shape
    Xp = torch.zeros(hin.shape[0], ops.pad4(in_i)); Xp[:, :in_i] = hin
    Y = torch.zeros(hin.shape[0], ops.pad4(out_i), device=DEV)
    ops.linear_act(Xp.to(DEV), wsplit[i], lins[i].bias, out_i, in_i, op=1, beta=100.0, oscale=1.0, out=Y[:, :out_i])
    print(f"layer {i} ({in_i}->{out_i}) isolated err {rel(Y[:, :out_i], hs[i]):.2e}  |h| max {float(hs[i].abs().max()):.3f} "
          f"|z| max {float(zs[i].abs().max()):.3f}", flush=True)
igr._debug_layers = []
f, ctx = igr.implicit_forward(net, x=x.to(DEV))
torch.cuda.synchronize()
for i, Y in enumerate(igr._debug_layers):
    ref = hs[i]
    if i == 3:
        ref = torch.cat([hs[3], x], -1) / math.sqrt(2)
    print(f"chain layer {i}: err {rel(Y[:, :ref.shape[1]], ref):.2e}", end="")
    if i == 3:
        print(f"  h part {rel(Y[:, :254], ref[:, :254]):.2e}  x part {rel(Y[:, 254:512], ref[:, 254:]):.2e}", end="")
    print(flush=True)
print("network f err", rel(f, zs[8]))
h8 = igr._debug_layers[7]
wl, bl = net.lin8.weight, net.lin8.bias
print("f from chain h8 by torch:", rel(h8 @ wl.t() + bl, zs[8]), " rowdots vs torch on same h8:", rel(f, h8 @ wl.t() + bl))
