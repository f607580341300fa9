"""Exploration: CUDA green contexts (driver API, via cuda-python) as a hard SM partition between the layer stream and the
coordinate / loss streams.  Step 1: create two green contexts (32 SMs | the rest), a stream in each, wrap them for torch,
run the same kernel in each and time it (a kernel that saturates the SMs should scale with the partition size)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuda.bindings import driver as cu


def chk(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


torch.cuda.init()
x = torch.zeros(1, device="cuda")          # primary context is current
dev = chk(cu.cuDeviceGet(0))
res = chk(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("device SMs", res.sm.smCount)
groups, nb, remaining = chk(cu.cuDevSmResourceSplitByCount(1, res, 0, 32))
print("groups", nb, "group0 SMs", groups[0].sm.smCount, "remaining SMs", remaining.sm.smCount)
desc_a = chk(cu.cuDevResourceGenerateDesc([groups[0]], 1))
desc_b = chk(cu.cuDevResourceGenerateDesc([remaining], 1))
g_a = chk(cu.cuGreenCtxCreate(desc_a, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
g_b = chk(cu.cuGreenCtxCreate(desc_b, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
s_a = chk(cu.cuGreenCtxStreamCreate(g_a, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
s_b = chk(cu.cuGreenCtxStreamCreate(g_b, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
ts_a, ts_b = torch.cuda.ExternalStream(int(s_a)), torch.cuda.ExternalStream(int(s_b))
print("streams", ts_a, ts_b)

a = torch.randn(8192, 8192, device="cuda")
b = torch.randn(8192, 8192, device="cuda")
torch.cuda.synchronize()


def timed(stream, n=3):
    with torch.cuda.stream(stream):
        for _ in range(2):
            a @ b
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(n):
            a @ b
        e.record(stream)
    e.synchronize()
    return s.elapsed_time(e) / n


print("matmul ms: default stream", round(timed(torch.cuda.current_stream()), 3), "| 32-SM green ctx", round(timed(ts_a), 3),
      "| rest green ctx", round(timed(ts_b), 3))
# graph capture on a green-context stream, replay on the default stream
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=ts_a):
    c = a @ b
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); g.replay(); e.record(); e.synchronize()
print("graph captured on the 32-SM stream, replayed on the default stream: ms", round(s.elapsed_time(e), 3))
with torch.cuda.stream(ts_a):
    s.record(ts_a); g.replay(); e.record(ts_a)
e.synchronize()
print("... replayed on the 32-SM stream: ms", round(s.elapsed_time(e), 3))
print("DONE")

# ---- step 2: the coordinate stage and the loss block of config 2 confined to the small partition ----
from point2cyl_b200 import ops, pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
B, N, K = 32, 8192, 8
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to("cuda").train()
batch = {k: v.to("cuda") for k, v in synthetic.s_cyl(B, N, K, seed=1234).items()}
start = [torch.randint(0, N, (B,)).cuda(), torch.randint(0, 512, (B,)).cuda()]
geo = pipeline.Geometry.empty(net, B, N, torch.device("cuda"))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def graph_time(fn, stream, tag):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g, stream=stream):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); print(tag, round(ts[len(ts) // 2], 4), "ms", flush=True)
    return g


for sms, grp in ((32, 0),):
    pass
graph_time(lambda: pipeline.geometry_forward(net, batch["pcs"], start, out=geo), None, "coordinate stage, whole GPU       ")
graph_time(lambda: pipeline.geometry_forward(net, batch["pcs"], start, out=geo), ts_a, "coordinate stage, 32-SM partition ")
with torch.no_grad():
    X_raw, W_raw = pipeline.backbone_forward(net, batch["pcs"], start, geo=geo)
lf = lambda: pipeline.loss_forward(batch["pcs"], X_raw, W_raw, batch["normals"], batch["inst"], batch["bb"], batch["axes"], batch["centers"])
graph_time(lf, None, "loss block, whole GPU             ")
graph_time(lf, ts_a, "loss block, 32-SM partition       ")
ops.set_sm_budget(116)
graph_time(lambda: pipeline.backbone_forward(net, batch["pcs"], None, geo=geo), ts_b, "layers, 116-SM partition          ")
ops.set_sm_budget(0)
print("DONE 2")
