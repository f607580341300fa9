"""Host-side cost of a pipelined step (enqueue time with the GPU idle / in steady state) for both pipeline classes."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point2cyl_b200 import synthetic, pin_batch
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
from point2cyl_b200.graph import DeepPipelinedForwardLoss, PipelinedForwardLoss
B, N, K = 32, 8192, 8
dev = torch.device("cuda")
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to(dev).train()
host = pin_batch(synthetic.s_cyl(B, N, K, seed=1234))
batch = {k: v.to(dev) for k, v in host.items()}
for cls in (PipelinedForwardLoss, DeepPipelinedForwardLoss):
    pipe = cls(net, batch)
    pipe.prime(None); pipe.step(None)
    for arg, tag in ((None, "resident"), (host, "host batch")):
        for _ in range(10): pipe.step(arg); pipe.join()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 300
        for _ in range(n):
            pipe.step(arg); pipe.join()
        t_host = (time.perf_counter() - t0) / n * 1e3        # host enqueue time per step (GPU may lag behind)
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / n * 1e3
        # pure host cost: three steps right after a synchronize (the FPS-start ring lets the host run 4 steps ahead)
        hs = []
        for _ in range(20):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(3):
                pipe.step(arg); pipe.join()
            hs.append((time.perf_counter() - t1) / 3 * 1e3)
        hs.sort()
        print(cls.__name__, tag, "pure host ms/step", round(hs[len(hs) // 2], 3), end=" | ")
        print("host enqueue ms/step", round(t_host, 3), "| wall ms/step incl. GPU", round(t_all, 3), flush=True)
