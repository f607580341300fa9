"""sa1 of config 2 (B=32 x 512 centres x 64 neighbours = 1,048,576 rows): the materialised first layer + 64 -> 64 layer
against p2c_sa_xyz_linear (first layer recomputed in the operand transform), with parts of the kernel switched off
(P2C_TC_DBG bit0: no first-layer math, bit1: no epilogue body, bit4 (16): no coordinate gather) to see which warp role
paces it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point2cyl_b200 import _lib, ops, synthetic
B, N, S, ns = 32, 8192, 512, 64
xyz = synthetic.s_cyl(B, N, 8, 1)["pcs"].cuda()
_, cxyz = ops.fps(xyz, S, torch.zeros(B, dtype=torch.long, device="cuda"))
gidx = ops.ball_query(0.2, ns, xyz, cxyz)
g = torch.Generator().manual_seed(0)
W0, b0 = torch.randn(64, 3, generator=g).cuda(), torch.randn(64, generator=g).cuda()
sc, sh = (torch.rand(64, generator=g) + 0.5).cuda(), torch.randn(64, generator=g).cuda()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(2000000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return min(ts[1:])


if os.environ.get("PAIR_EXP"):
    # sa1.1 alone: P2C_SA_PAIR=0 (channels-as-lanes kernel of linear_tc.cu) against =1 (rows-as-lanes, sa_stack_tc.cu MODE 1)
    W1, b1 = (torch.randn(64, 64, generator=g) / 8).cuda(), torch.randn(64, generator=g).cuda()
    st = torch.zeros(128, dtype=torch.float64, device="cuda")
    t = timed(lambda: ops.sa_xyz_linear(xyz, cxyz, gidx, W0, b0, W1, b1, scale0=sc, shift0=sh, stats=st, pool_group=0, want_y=True), reps=8)
    print(f"sa_xyz_linear 64 -> 64: {t:.1f} us  (P2C_SA_PAIR={os.environ.get('P2C_SA_PAIR')})")
    sys.exit(0)
for N1, pool in ((64, 0), (128, 64)):
    W1, b1 = (torch.randn(N1, 64, generator=g) / 8).cuda(), torch.randn(N1, generator=g).cuda()
    st = torch.zeros(2 * N1, dtype=torch.float64, device="cuda")
    Y0 = ops.sa_first_layer(xyz, cxyz, gidx, None, W0, b0, None)
    print(f"64 -> {N1} pool {pool}: sa_first_layer {timed(lambda: ops.sa_first_layer(xyz, cxyz, gidx, None, W0, b0, None)):.1f} us")
    for mode in os.environ.get("MODES", "0 1 2 16 17 19").split():
        os.environ["P2C_TC_DBG"] = mode
        t_lin = timed(lambda: ops.linear(Y0, W1, b1, in_scale=sc, in_shift=sh, stats=st, pool_group=pool, want_y=(pool == 0), precision=1))
        t_xyz = timed(lambda: ops.sa_xyz_linear(xyz, cxyz, gidx, W0, b0, W1, b1, scale0=sc, shift0=sh, stats=st, pool_group=pool, want_y=(pool == 0)))
        print(f"   P2C_TC_DBG={mode:>2}: linear {t_lin:6.1f} us   sa_xyz_linear {t_xyz:6.1f} us", flush=True)
    os.environ["P2C_TC_DBG"] = "0"
