"""Two warm-up training steps, then ONE step inside cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
from point2cyl_b200.train import Trainer

B, N, K = 32, 8192, 8
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).cuda().train()
batch = {k: v.cuda() for k, v in synthetic.s_cyl(B, N, K, seed=1234).items()}
tr = Trainer(net)
for _ in range(2):
    tr.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
