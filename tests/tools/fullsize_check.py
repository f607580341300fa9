"""GPU box: FPS, ball query (both levels) and 3-NN of the kernels against the C oracle at BASELINE.json's FULL sizes -
the whole config-2 batch (B=32 x N=8192, S-cyl) and 16 clouds of the stress configuration (N=32768, S-uniform) - bit for
bit.  Not collected by pytest yet: written at the end of round 1 without GPU time left to run it; promote it to
tests/test_gpu_parity.py once it has been seen green.

    python tests/tools/fullsize_check.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from oracle import c_oracle as corc  # noqa: E402
from point2cyl_b200 import ops, synthetic  # noqa: E402

DEV = "cuda:0"


def check(B, N, kind):
    xyz = synthetic.s_cyl(B, N, 8, 1234)["pcs"] if kind == "cyl" else synthetic.s_uniform(B, N, 5)
    g = torch.Generator().manual_seed(5)
    s1 = torch.randint(0, N, (B,), generator=g)
    s2 = torch.randint(0, 512, (B,), generator=g)
    xg = xyz.to(DEV)
    t0 = time.perf_counter()
    idx1, c1 = ops.fps(xg, 512, s1.to(DEV))
    grp1 = ops.ball_query(0.2, 64, xg, c1)
    idx2, c2 = ops.fps(c1, 128, s2.to(DEV))
    grp2 = ops.ball_query(0.4, 64, c1, c2)
    feats = torch.randn(B, 512, 16, generator=g)
    out, nidx, w = ops.three_nn_interp(xg, c1, feats.reshape(-1, 16).to(DEV), want_idx=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r1 = corc.farthest_point_sample(xyz, 512, s1)
    ok = {"fps1": torch.equal(idx1.cpu(), r1)}
    c1h = c1.cpu()
    ok["ball1"] = torch.equal(grp1.cpu(), corc.query_ball_point(0.2, 64, xyz, c1h))
    r2 = corc.farthest_point_sample(c1h, 128, s2)
    ok["fps2"] = torch.equal(idx2.cpu(), r2)
    ok["ball2"] = torch.equal(grp2.cpu(), corc.query_ball_point(0.4, 64, c1h, c2.cpu()))
    ridx, rw, _ = corc.three_nn(xyz, c1h)
    ok["nn_idx"] = torch.equal(nidx.cpu().reshape(ridx.shape), ridx)
    ok["nn_w"] = torch.equal(w.cpu().reshape(rw.shape), rw)
    print(f"B={B} N={N} {kind}: kernels {t1 - t0:.2f}s, C oracle {time.perf_counter() - t1:.1f}s ->", ok, flush=True)
    return all(ok.values())


if __name__ == "__main__":
    good = check(32, 8192, "cyl") & check(16, 32768, "uniform")
    print("FULL-SIZE PARITY", "OK" if good else "MISMATCH")
    sys.exit(0 if good else 1)
