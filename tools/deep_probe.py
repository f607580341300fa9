"""Where the three-stage pipelined step spends its time: the same DeepPipelinedForwardLoss with the coordinate stage and / or
the loss stage switched off (their graphs are simply not replayed; results are then stale - timing only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point2cyl_b200 import synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
from point2cyl_b200.graph import DeepPipelinedForwardLoss
B, N, K = 32, 8192, 8
dev = torch.device("cuda")
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to(dev).train()
batch = {k: v.to(dev) for k, v in synthetic.s_cyl(B, N, K, seed=1234).items()}
pipe = DeepPipelinedForwardLoss(net, batch)
pipe.prime(None); pipe.step(None)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


class Off:
    def replay(self):
        pass


def run(tag):
    for _ in range(5):
        pipe.step(None); pipe.join()
    ts = []
    for _ in range(30):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); pipe.step(None); pipe.join(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); print(tag, round(ts[len(ts) // 2], 4), "ms", flush=True)


geo, loss = list(pipe.g_geo), list(pipe.g_loss)
run("all three stages        ")
pipe.g_loss = [Off()] * 3
run("no loss stage           ")
pipe.g_loss = loss; pipe.g_geo = [Off()] * 3
run("no coordinate stage     ")
pipe.g_loss = [Off()] * 3
run("layers only (116 SMs)   ")
