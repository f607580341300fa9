"""Timeline of the warp roles of CTA 0 in the tcgen05 layer kernel: tc_timeline.py M K N pool want_y stats"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops
M, K, Nn, pool, want_y, use_stats = (int(v) for v in sys.argv[1:7])
ld = ops.pad4(K)
X = torch.randn(M, ld, device="cuda"); W = torch.randn(Nn, K, device="cuda") / K ** 0.5
b = torch.randn(Nn, device="cuda"); sc = torch.rand(K, device="cuda") + 0.5; sh = torch.randn(K, device="cuda")
dbg = torch.zeros(4 * 512 + 256 * 4, dtype=torch.int64, device="cuda")
def run():
    stats = torch.zeros(2 * Nn, dtype=torch.float64, device="cuda") if use_stats else None
    ops.linear(X, W, b, K=K, in_scale=sc, in_shift=sh, stats=stats, pool_group=pool, want_y=bool(want_y), precision=1)
run(); run(); torch.cuda.synchronize()
_lib.load().p2c_debug_set_timeline(dbg.data_ptr())
run(); torch.cuda.synchronize()
_lib.load().p2c_debug_set_timeline(None)
cta = dbg.cpu()[2048:].reshape(256, 4)
d = dbg.cpu()[:2048].reshape(4, 256, 2)
t0 = int(d[:, :, 1][d[:, :, 1] > 0].min())
names = ["mma", "xform", "epi", "prod"]
ev = []
for r in range(4):
    for i in range(256):
        if d[r, i, 1] > 0:
            ev.append((int(d[r, i, 1]) - t0, names[r], int(d[r, i, 0])))
ev.sort()
lim = int(sys.argv[7]) if len(sys.argv) > 7 else 150
for t, n, tag in ev:
    if tag >= 9000 or lim > 0:
        print(f"{t:8d} ns  {n:6s} tile {tag // 100:3d} ev {tag % 100:2d}")
    lim -= 1
last = {}
for t, n, tag in ev:
    last[n] = (t, tag)
print("last events", last)

live = cta[cta[:, 0] > 0]
base = int(live[:, 0].min())
import statistics
ent = [int(x) - base for x in live[:, 0]]
pro = [int(x) - base for x in live[:, 1]]
end = [int(x) - base for x in live[:, 2]]
print(f"CTAs {len(ent)}: entry min/max {min(ent)}/{max(ent)} ns; prologue-done min/med/max {min(pro)}/{statistics.median(pro)}/{max(pro)}; "
      f"roles-done min/med/max {min(end)}/{statistics.median(end)}/{max(end)}")
