import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point2cyl_b200 import ops, pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
B, N, K = 32, 8192, 8
dev = torch.device("cuda")
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to(dev).train()
batch = {k: v.to(dev) for k, v in synthetic.s_cyl(B, N, K, seed=1234).items()}
start = [torch.randint(0, N, (B,)).to(dev), torch.randint(0, 512, (B,)).to(dev)]
geo = pipeline.geometry_forward(net, batch["pcs"], start)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for budget in (148, 132, 116, 100):
    ops.set_sm_budget(budget)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2): pipeline.backbone_forward(net, batch["pcs"], None, geo=geo)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        pipeline.backbone_forward(net, batch["pcs"], None, geo=geo)
    ops.set_sm_budget(0)
    for _ in range(3): g.replay()
    ts = []
    for _ in range(20):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); print("budget", budget, "backbone-only graph ms", round(ts[len(ts)//2], 4), flush=True)
