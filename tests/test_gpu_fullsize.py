"""GPU parity at BASELINE.json's FULL sizes and in the headline's own mode (train-mode BatchNorm), plus the float64
adjudication of every tolerance that is looser than north_star's 1e-4 (tests/adjudication.py).

  * point operators bit for bit against the C oracle over the whole config-2 batch (32 x 8192, S-cyl) and 16 clouds of
    the stress configuration (32768 points, S-uniform): FPS and ball query at both levels, 3-NN indices and weights;
  * config 2 (B=32, N=8192, K=8), train-mode AND running-statistics BatchNorm: every index tensor of the backbone
    exact, X_raw / W_raw / BatchNorm running statistics / six losses / fitted axes / centres against the oracle at 1e-4;
  * golden batches (B <= 2: the ill-conditioned case) and their gradients: |kernel - fp64| <= max(1e-4, c*|ref32 - fp64|).

Each test appends its measured errors to gpurun_out/parity_report.jsonl when that directory exists.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as corc
from oracle import p2c_oracle as orc
from point2cyl_b200 import ops, synthetic
from tests import adjudication as adj

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def report(name, payload):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **payload}) + "\n")


@pytest.mark.parametrize("B,N,kind", [(32, 8192, "cyl"), (16, 32768, "uniform")])
def test_pointops_fullsize_bit_exact(B, N, kind):
    """the whole config-2 batch and 16 stress clouds, bit for bit against oracle/p2c_oracle_c.c
    (itself bit-pinned to the reference goldens, tests/test_oracle_c.py).  pointnet_util.py:63-107, 298-308."""
    corc.build()                      # gcc on the box if the prebuilt checker did not travel
    xyz = synthetic.s_cyl(B, N, 8, 1234)["pcs"] if kind == "cyl" else synthetic.s_uniform(B, N, 5)
    g = torch.Generator().manual_seed(5)
    s1 = torch.randint(0, N, (B,), generator=g)
    s2 = torch.randint(0, 512, (B,), generator=g)
    xg = xyz.to(DEV)
    idx1, c1 = ops.fps(xg, 512, s1.to(DEV))
    grp1 = ops.ball_query(0.2, 64, xg, c1)
    idx2, c2 = ops.fps(c1, 128, s2.to(DEV))
    grp2 = ops.ball_query(0.4, 64, c1, c2)
    feats = torch.randn(B, 512, 16, generator=g)
    _, nidx, w = ops.three_nn_interp(xg, c1, feats.reshape(-1, 16).to(DEV), want_idx=True)
    _, nidx2, w2 = ops.three_nn_interp(c1, c2, feats[:, :128].reshape(-1, 16).contiguous().to(DEV), want_idx=True)
    c1h, c2h = c1.cpu(), c2.cpu()
    assert torch.equal(idx1.cpu(), corc.farthest_point_sample(xyz, 512, s1))
    assert torch.equal(c1h, orc.gather_points(xyz, idx1.cpu()))
    assert torch.equal(grp1.cpu(), corc.query_ball_point(0.2, 64, xyz, c1h))
    assert torch.equal(idx2.cpu(), corc.farthest_point_sample(c1h, 128, s2))
    assert torch.equal(grp2.cpu(), corc.query_ball_point(0.4, 64, c1h, c2h))
    ridx, rw, _ = corc.three_nn(xyz, c1h)
    assert torch.equal(nidx.cpu().reshape(ridx.shape), ridx)
    assert torch.equal(w.cpu().reshape(rw.shape), rw)
    ridx2, rw2, _ = corc.three_nn(c1h, c2h)
    assert torch.equal(nidx2.cpu().reshape(ridx2.shape), ridx2)
    assert torch.equal(w2.cpu().reshape(rw2.shape), rw2)


def _assert_bars(b, name, slack, keys=None, tol=TOL, propagated=0.0):
    """e_kern <= max(tol, slack * e_ref[, propagated]).  `propagated`: for quantities DERIVED from the network outputs
    (losses, fitted axes, centres) the error the outputs themselves carry - the loss block on given outputs is held to
    1e-4 separately (test_loss_golden), so a derived quantity only has to not amplify what it is fed."""
    bad = []
    for k, v in b.items():
        if keys is not None and not any(k.startswith(p) for p in keys):
            continue
        if v["e_kern"] is None:
            ok = v["e_direct"] <= tol
        else:
            extra = 0.0 if k in ("X_raw", "W_raw") else propagated
            ok = v["e_direct"] <= tol or v["e_kern"] <= max(tol, slack * v["e_ref"], extra)
        if not ok:
            bad.append((k, v))
    assert not bad, (name, bad)


@pytest.mark.parametrize("training", [True, False])
def test_config2_full_batch_vs_oracle(training):
    """VERDICT r1 item 1b.  BASELINE.json configs[1] at full size in the mode bench.py times (train-mode BatchNorm,
    dropout on) and in eval.py's mode: all discrete choices exact, all floats within 1e-4 of the oracle -
    models/pointnet_extrusion.py:37-66, train_Point2Cyl_without_sketch.py:244-353."""
    B, N, K = 32, 8192, 8
    starts = (torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(1)),
              torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(2)))
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(9)) > 0.5).float() * 2.0
    sd = orc.init_state_dict((3, 2 * K), 0)
    case = adj.run_case(B, N, K, 1234, training, starts, mask, grads=False, sd=sd, want64=True,
                        force_kernel_choices=True)
    ko, kt, net, _ = case["kern"]
    r32 = case["ref32"][0]
    # ---- discrete choices against the reference algorithm run on the same inputs, all 32 clouds ----
    ch, pcs = case["forced"], case["data"]["pcs"]
    f1 = orc.farthest_point_sample(pcs, 512, starts[0])
    assert torch.equal(ch["sa1.fps_idx"], f1)
    l1_xyz = orc.gather_points(pcs, f1)
    assert torch.equal(kt["l1_xyz"].cpu(), l1_xyz)
    assert torch.equal(ch["sa1.group_idx"], orc.query_ball_point(0.2, 64, pcs, l1_xyz))
    f2 = orc.farthest_point_sample(l1_xyz, 128, starts[1])
    assert torch.equal(ch["sa2.fps_idx"], f2)
    l2_xyz = orc.gather_points(l1_xyz, f2)
    assert torch.equal(ch["sa2.group_idx"], orc.query_ball_point(0.4, 64, l1_xyz, l2_xyz))
    ties = {}
    for name, q, s_ in (("fp1", pcs, l1_xyz), ("fp2", l1_xyz, l2_xyz)):
        d, order = orc.square_distance(q, s_).sort(dim=-1)
        diff = (order[:, :, :3] != ch[name + ".nn_idx"]).nonzero()
        # The reference takes the first three entries of an UNSTABLE sort (pointnet_util.py:301-303): at exactly equal
        # distances its pick is implementation-defined; the kernels (and the C oracle) take the lowest index.  Any
        # difference must be such a tie, and the neighbour DISTANCES must still agree bit for bit.
        for bb, n, j in diff.tolist():
            mine = int(ch[name + ".nn_idx"][bb, n, j])
            pos = int((order[bb, n] == mine).nonzero()[0])
            assert float(d[bb, n, pos]) == float(d[bb, n, j]), (name, bb, n, j)
        ties[name] = len(diff)
        assert len(diff) <= 16, (name, len(diff))
        w_ref = 1.0 / (d[:, :, :3] + 1e-8)
        w_ref = w_ref / w_ref.sum(dim=2, keepdim=True)
        assert torch.equal(ch[name + ".nn_w"], w_ref), name      # weights depend on the distances only: exact
    assert torch.equal(ko["matching_indices"].cpu(), r32["matching_indices"])
    assert torch.equal(ko["mask"].cpu(), r32["mask"])
    b = adj.bars(case)
    w = adj.worst(b)
    report(f"config2_full_batch[{'train' if training else 'eval'}]",
           {"worst": w, "nn_ties_resolved_differently_by_torch_sort": ties, "bars": {k: v for k, v in b.items() if not k.startswith("stat:")},
            "worst_stat": adj.worst(b, "stat:")})
    # the bar north_star states, against the float32 oracle directly
    direct_bad = {k: v["e_direct"] for k, v in b.items() if v["e_direct"] > TOL}
    assert not direct_bad, direct_bad


GOLDEN_BACKBONE = ["backbone_b2_n1024_k4.npz", "backbone_b1_n1024_k4.npz"]
# How much less exact than the reference's own float32 run the kernels may be, per quantity, before a test fails.
# Measured (tests/tools/precision_probe.py, profiles/r2a_precision_probe.json): one 3xTF32 layer on the tensor cores is
# 0.8e-6 (K=64) ... 1.3e-6 (K=128) ... 9.6e-6 (K=1280) rms from the exact product where an fp32 FMA chain is
# 1.0e-7 ... 4.6e-7: the operand split itself is fp32-grade (4e-7 with exact accumulation), the difference is the
# tensor core's TRUNCATING fp32 accumulate (48 accumulate steps at K=128).  That is a factor 6-20 per layer; measured
# end to end on these ill-conditioned tiny batches the kernels are 11x (B=2) / 3x (B=1) the reference's own distance
# from the exact result, and 1.8x at the headline batch (B=32: 5.7e-5 vs 3.1e-5, test_config2_full_batch_vs_oracle,
# where the plain 1e-4 bar holds without any adjudication).
SLACK_FWD = 16.0


@pytest.mark.parametrize("name", GOLDEN_BACKBONE)
def test_golden_train_mode_fp64_adjudicated(golden_dir, name):
    """the ill-conditioned tiny-batch train-mode goldens.  The reference's OWN float32 outputs
    are e_ref from the exact (float64) value; the kernels must be within max(1e-4, SLACK_FWD * e_ref) of it."""
    g = np.load(os.path.join(golden_dir, name))
    B, N, K, seed = (int(v) for v in g["meta"])
    starts = (torch.from_numpy(g["train_s1"]), torch.from_numpy(g["train_s2"]))
    case = adj.run_case(B, N, K, seed, True, starts, None, grads=False)
    # the float32 oracle IS the reference here (bit-equal to the golden)
    assert adj.rel_max(case["ref32"][0]["X_raw"], g["train_X"]) <= 1e-6
    b = adj.bars(case)
    report(f"golden_train_fp64[{name}]", {"worst": adj.worst(b), "bars": {k: b[k] for k in ("X_raw", "W_raw", "total")}})
    _assert_bars(b, name, SLACK_FWD, propagated=max(b["X_raw"]["e_kern"], b["W_raw"]["e_kern"]))


SLACK_GRAD = 16.0      # same factor as the forward (measured worst 8.3x, median 3.9x)


@pytest.mark.parametrize("name,training", [("train_b2_n1024_k4.npz", True), ("train_bneval_b2_n1024_k4.npz", False)])
def test_golden_gradients_fp64_adjudicated(golden_dir, name, training):
    """The end-to-end gradient bars (tests/test_gpu_backward.py: 2e-1 train-mode, 5e-3 running statistics) next to
    their float64 bound: per parameter, relative L2 of (kernel - fp64) <= max(BAR, SLACK_GRAD * rel L2 of
    (reference fp32 - fp64)), BAR = 1e-4 in train mode.  With running statistics the reference is ~1e-6 from exact
    and the kernels' error is isolated ReLU / arg-max flips of 3xTF32-vs-fp32 rounding: BAR = 5e-3 there (stated)."""
    from tests.test_oracle_golden import live_keys
    g = np.load(os.path.join(golden_dir, name))
    B, N, K, seed = (int(v) for v in g["meta"])
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0
    starts = (torch.from_numpy(g["s1"]), torch.from_numpy(g["s2"]))
    case = adj.run_case(B, N, K, seed, training, starts, mask, grads=True)
    assert torch.equal(case["ref32"][0]["matching_indices"], case["ref64"][0]["matching_indices"])
    b = adj.bars(case, with_stats=False)
    live = ["grad:" + k for k in live_keys(g, training)]
    gb = {k: v for k, v in b.items() if k in live}
    assert len(gb) >= 40
    report(f"golden_grad_fp64[{name}]", {"worst": adj.worst(gb), "n": len(gb),
                                        "median_e_kern": float(np.median([v["e_kern"] for v in gb.values()])),
                                        "median_e_ref": float(np.median([v["e_ref"] for v in gb.values()]))})
    _assert_bars(gb, name, SLACK_GRAD, tol=TOL if training else 5e-3)
