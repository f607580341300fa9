"""torchrun --nproc-per-node 2 tools/check_dp_train.py : data-parallel Trainer over NCCL.
Checks that (1) the all-reduced flat gradient equals the mean of the ranks' local gradients, (2) parameters stay
bit-identical across ranks after optimiser steps, (3) the loss decreases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from point2cyl_b200 import dist as pd
from point2cyl_b200 import synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
from point2cyl_b200.train import Trainer

rank, world, local = pd.init("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
B, N, K = 8, 4096, 8
glob = synthetic.s_cyl(B * world, N, K, seed=5)
mine = {k: v.to(dev) for k, v in pd.shard_batch(glob, rank, world).items()}
torch.manual_seed(0)                       # same initial weights on every rank
net = backbone(output_sizes=[3, 2 * K]).to(dev).train()
tr = Trainer(net, lr=1e-3)
starts = (torch.zeros(B, dtype=torch.long, device=dev), torch.zeros(B, dtype=torch.long, device=dev))
torch.manual_seed(100 + rank)              # dropout masks differ per rank, like independent replicas
tr.forward_backward(mine, fps_start=starts)
local_grad = tr.flat_grad.clone()
gathered = [torch.empty_like(local_grad) for _ in range(world)]
dist.all_gather(gathered, local_grad)
mean_grad = torch.stack(gathered).mean(0)
dist.all_reduce(tr.flat_grad, op=dist.ReduceOp.SUM)
err = float((tr.flat_grad / world - mean_grad).abs().max() / mean_grad.abs().max())
assert err < 1e-6, err
losses = []
for _ in range(8):
    losses.append(float(tr.step(mine, fps_start=starts)["total"]))
p = tr.flat_param.clone()
ps = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(ps, p)
assert all(torch.equal(ps[0], q) for q in ps), "parameters diverged across ranks"
assert losses[-1] < losses[0], losses
if rank == 0:
    print(f"dp{world} ok: grad all-reduce err {err:.1e}, params identical on all ranks, loss {losses[0]:.4f} -> {losses[-1]:.4f}")
dist.destroy_process_group()
