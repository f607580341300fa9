"""Per-kernel time of one Trainer step (forward + loss + backward + Adam) at B=32 x N=8192, K=8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, synthetic, pipeline
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
from point2cyl_b200.train import Trainer

B, N, K = int(os.environ.get("B", 32)), 8192, 8
dev = "cuda"
torch.manual_seed(0)
net = backbone(output_sizes=[3, 2 * K]).to(dev).train()
batch = {k: v.to(dev) for k, v in synthetic.s_cyl(B, N, K, seed=1234).items()}
tr = Trainer(net)
for _ in range(3):
    out = tr.step(batch)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    tr.step(batch)
e.record(); e.synchronize()
print(f"eager train step: {s.elapsed_time(e) / 5:.3f} ms  loss {float(out['total']):.4f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
_lib.profile_start()
tr.step(batch)
agg = {}
for name, tag, t in _lib.profile_stop():
    key = (name, tag.split(".")[0] if not tag.startswith("bwd") else tag)
    a = agg.setdefault(key, [0.0, 0]); a[0] += t; a[1] += 1
tot = sum(v[0] for v in agg.values())
byname = {}
for (name, tag), (t, n) in agg.items():
    b = byname.setdefault(("bwd " if tag.startswith("bwd") else "fwd ") + name, [0.0, 0]); b[0] += t; b[1] += n
print(f"sum of kernel times {tot:.3f} ms")
for k, (t, n) in sorted(byname.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:34s} {t:8.3f} ms  x{n}")
print("backward by stage / kernel:")
for (name, tag), (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {tag:10s} {name:26s} {t:8.3f} ms x{n}")
