import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from point2cyl_b200 import ops
g = torch.Generator().manual_seed(0)
for (B, N, S, ns, C, D) in [(2, 1024, 512, 64, 64, 0), (2, 512, 128, 64, 128, 128), (1, 1024, 512, 64, 64, 0)]:
    xyz = torch.rand(B, N, 3, generator=g); new_xyz = torch.rand(B, S, 3, generator=g)
    idx = torch.randint(0, N, (B, S, ns), generator=g)
    W = torch.randn(C, 3 + D, generator=g); bias = torch.randn(C, generator=g) + 3.0
    feats = torch.randn(B * N, D, generator=g) if D else None
    Qf = (feats @ W[:, 3:].T).cuda() if D else None
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    Y = ops.sa_first_layer(xyz.cuda(), new_xyz.cuda(), idx.cuda(), Qf, W.cuda(), bias.cuda(), stats)
    rel = (torch.gather(xyz.unsqueeze(1).expand(B, S, N, 3), 2, idx.unsqueeze(-1).expand(B, S, ns, 3)) - new_xyz.unsqueeze(2)).reshape(-1, 3).double()
    ref = rel @ W[:, :3].double().T + bias.double()
    if D:
        rows = (torch.arange(B).view(B, 1, 1) * N + idx).reshape(-1)
        ref = ref + (feats.double() @ W[:, 3:].double().T)[rows]
    print((B, N, S, ns, C, D), "Y err %.2e" % float((Y.cpu().double() - ref).abs().max()),
          "s1 relerr %.2e" % float(((stats[:C].cpu() - ref.sum(0)).abs() / ref.sum(0).abs()).max()),
          "s2 relerr %.2e" % float(((stats[C:].cpu() - (ref ** 2).sum(0)).abs() / (ref ** 2).sum(0)).max()),
          "Ysum-vs-stats %.2e" % float(((stats[:C].cpu() - Y.cpu().double().sum(0)).abs() / ref.sum(0).abs()).max()))
