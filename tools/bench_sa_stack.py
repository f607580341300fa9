"""sa1 of config 2 (B=32 x N=8192, S=512, ns=64, 3 -> 64 -> 64 -> 128) in eval mode: the one-kernel level
(p2c_sa_stack_fused) against the per-layer kernels (p2c_sa_xyz_linear + p2c_linear + p2c_pool_bn_relu); L2 flushed
between timed runs.  Also the whole eval-mode backbone forward both ways."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point2cyl_b200 import _lib, ops, pipeline, synthetic  # noqa: E402
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone  # noqa: E402

B, N, K = 32, 8192, 8
dev = torch.device("cuda")
data = synthetic.s_cyl(B, N, K, seed=1234)
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to(dev).eval()
pcs = data["pcs"].to(dev)
start = [torch.randint(0, N, (B,)).to(dev), torch.randint(0, 512, (B,)).to(dev)]
geo = pipeline.geometry_forward(net, pcs, start, moments=False)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


out = {}
with torch.no_grad():
    for name, flag in (("fused", True), ("per_layer", False)):
        pipeline.stack_fused_enabled = flag
        out[f"sa1_{name}_ms"] = timed(lambda: pipeline.set_abstraction(net.sa1, geo.xyz, None, None,
                                                                      geo=(geo.fps1, geo.l1_xyz, geo.gidx1)))
        out[f"backbone_eval_{name}_ms"] = timed(lambda: pipeline.backbone_forward(net, pcs, start, geo=geo))
        _lib.profile_start()
        pipeline.set_abstraction(net.sa1, geo.xyz, None, None, geo=(geo.fps1, geo.l1_xyz, geo.gidx1))
        out[f"sa1_{name}_kernels"] = [(n, round(t * 1e3, 1)) for n, _, t in _lib.profile_stop()]
    # the whole eval-mode backbone as ONE CUDA graph each way (eager launches above are host bound)
    for name, flag in (("fused", True), ("per_layer", False)):
        pipeline.stack_fused_enabled = flag
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                pipeline.backbone_forward(net, pcs, start, geo=geo)
        torch.cuda.current_stream().wait_stream(side)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            res = pipeline.backbone_forward(net, pcs, start, geo=geo)
        out[f"backbone_eval_graph_{name}_ms"] = timed(gr.replay)
pipeline.stack_fused_enabled = True
rows = B * 512 * 64
flop = 3 * 2.0 * rows * (64 * 64 + 64 * 128)
out["fused_tf32_mma_tflops"] = flop / (out["sa1_fused_ms"] / 1e3) / 1e12
print(json.dumps(out))
