"""Run a few eager forward+loss steps at the benchmark configuration (for ncu captures): one_step.py [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from point2cyl_b200 import pipeline, synthetic
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net = bench.make_net("cuda")
batch = {k: v.cuda() for k, v in synthetic.s_cyl(bench.B_PER_GPU, bench.N_POINTS, bench.K_INST, seed=1234).items()}
with torch.no_grad():
    for _ in range(steps):
        out = pipeline.forward_loss(net, batch)
torch.cuda.synchronize()
print("loss", float(out["total"]))
