"""GPU tests of entry points that had none and of this round's host-side changes:
index_points / p2c_gather_rows, PointNetSetAbstractionMsg (forward and autograd), the loss block at the stress
configuration's K = 16, the in-kernel Philox dropout mask, graph re-capture after update_momentum."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import p2c_oracle as orc
from point2cyl_b200 import _lib, ops, pipeline, synthetic
from point2cyl_b200.dropin.models import pointnet_util as dpu
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().cpu().double().reshape(-1)
    b = torch.as_tensor(b).detach().cpu().double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ---- a3: index_points (models/pointnet_util.py:43-60) ------------------------------------------------------------

@pytest.mark.parametrize("B,N,C,shape", [(3, 700, 3, (40,)), (2, 1024, 16, (50, 8)), (2, 333, 131, (7, 5)),
                                         (1, 64, 1, (64,)), (4, 8192, 3, (512, 64))])
def test_index_points_matches_reference_gather(B, N, C, shape):
    g = torch.Generator().manual_seed(N + C)
    pts = torch.randn(B, N, C, generator=g)
    idx = torch.randint(0, N, (B,) + shape, generator=g)
    idx[0].reshape(-1)[:2] = torch.tensor([0, N - 1])            # both ends of the range
    ref = orc.gather_points(pts, idx)
    got = dpu.index_points(pts.to(DEV), idx.to(DEV))
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)
    got2 = ops.gather_rows(pts.to(DEV), idx.to(DEV))
    assert torch.equal(got2.cpu(), ref)


def test_index_points_strided_input():
    """a channel slice of a wider buffer and a permuted (B,C,N)->(B,N,C) view are valid `points` arguments"""
    g = torch.Generator().manual_seed(3)
    wide = torch.randn(2, 500, 24, generator=g)
    idx = torch.randint(0, 500, (2, 33, 4), generator=g)
    got = dpu.index_points(wide.to(DEV)[:, :, 4:20], idx.to(DEV))
    assert torch.equal(got.cpu(), orc.gather_points(wide[:, :, 4:20].contiguous(), idx))
    cf = torch.randn(2, 10, 500, generator=g)
    got = dpu.index_points(cf.to(DEV).permute(0, 2, 1), idx.to(DEV))
    assert torch.equal(got.cpu(), orc.gather_points(cf.permute(0, 2, 1).contiguous(), idx))


# ---- PointNetSetAbstractionMsg (models/pointnet_util.py:210-267) --------------------------------------------------

def msg_reference(mod_cpu, xyz_cf, points_cf, start):
    """The reference forward (:229-267) restated on torch CPU with the oracle's point operators."""
    xyz = xyz_cf.permute(0, 2, 1)
    points = points_cf.permute(0, 2, 1) if points_cf is not None else None
    B, N, C = xyz.shape
    S = mod_cpu.npoint
    new_xyz = orc.gather_points(xyz, orc.farthest_point_sample(xyz, S, start))
    outs = []
    for i, radius in enumerate(mod_cpu.radius_list):
        gidx = orc.query_ball_point(radius, mod_cpu.nsample_list[i], xyz, new_xyz)
        gx = orc.gather_points(xyz, gidx) - new_xyz.view(B, S, 1, C)
        gp = torch.cat([orc.gather_points(points, gidx), gx], dim=-1) if points is not None else gx
        h = gp.permute(0, 3, 2, 1)
        for conv, bn in zip(mod_cpu.conv_blocks[i], mod_cpu.bn_blocks[i]):
            h = F.relu(bn(conv(h)))
        outs.append(h.max(dim=2)[0])
    return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=1)


@pytest.mark.parametrize("D", [0, 6])
@pytest.mark.parametrize("training", [False, True])
def test_set_abstraction_msg_forward(D, training):
    torch.manual_seed(5)
    B, N = 2, 600
    mk = lambda: dpu.PointNetSetAbstractionMsg(48, [0.2, 0.4], [16, 32], D, [[16, 32], [24, 40]])
    mod = mk()
    for m in mod.modules():                                   # non-trivial running statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.uniform_(-0.1, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    ref_mod = mk()
    ref_mod.load_state_dict(mod.state_dict())
    mod, ref_mod = mod.to(DEV).train(training), ref_mod.train(training)
    xyz = synthetic.s_uniform(B, N, 8).permute(0, 2, 1).contiguous()
    pts = torch.randn(B, D, N) if D else None
    start = torch.tensor([3, 77])
    real = pipeline.draw_fps_start
    pipeline.draw_fps_start = lambda B_, N_, dev: start.to(dev)
    try:
        with torch.no_grad():
            new_xyz, out = mod(xyz.to(DEV), None if pts is None else pts.to(DEV))
            rx, ro = msg_reference(ref_mod, xyz, pts, start)
    finally:
        pipeline.draw_fps_start = real
    assert out.shape == ro.shape == (B, 32 + 40, 48)
    assert torch.equal(new_xyz.cpu(), rx)
    assert rel_err(out, ro) <= TOL
    if training:
        for (k, a), (_, b) in zip(mod.state_dict().items(), ref_mod.state_dict().items()):
            if "running" in k:
                assert rel_err(a, b) <= TOL, k


def test_set_abstraction_msg_autograd():
    """the Msg module had no grad_fn.  Parameter and feature gradients against torch autograd of the
    restated reference forward (BatchNorm on running statistics: the well-conditioned case)."""
    torch.manual_seed(6)
    B, N, D = 2, 500, 5
    mk = lambda: dpu.PointNetSetAbstractionMsg(40, [0.25, 0.5], [16, 32], D, [[32, 64], [64]])   # backward widths: 32 / 64 / k*128
    mod, ref_mod = mk(), mk()
    ref_mod.load_state_dict(mod.state_dict())
    mod, ref_mod = mod.to(DEV).eval(), ref_mod.eval()
    xyz = synthetic.s_uniform(B, N, 9).permute(0, 2, 1).contiguous()
    pts = torch.randn(B, D, N)
    start = torch.tensor([1, 2])
    p_dev = pts.to(DEV).requires_grad_(True)
    p_ref = pts.clone().requires_grad_(True)
    real = pipeline.draw_fps_start
    pipeline.draw_fps_start = lambda B_, N_, dev: start.to(dev)
    try:
        _, out = mod(xyz.to(DEV), p_dev)
    finally:
        pipeline.draw_fps_start = real
    assert out.grad_fn is not None
    _, ro = msg_reference(ref_mod, xyz, p_ref, start)
    assert rel_err(out, ro) <= TOL
    w = torch.randn(ro.shape, generator=torch.Generator().manual_seed(1))
    (out * w.to(DEV)).sum().backward()
    (ro * w).sum().backward()
    assert rel_l2(p_dev.grad, p_ref.grad) <= 5e-3
    for (k, a), (_, b) in zip(mod.named_parameters(), ref_mod.named_parameters()):
        assert a.grad is not None, k
        assert rel_l2(a.grad, b.grad) <= 5e-3, k


def test_module_backward_does_not_clobber_the_incoming_gradient():
    """the in-place backward kernels must not write into the gradient tensor autograd hands in."""
    torch.manual_seed(2)
    fp = dpu.PointNetFeaturePropagation(16 + 8, [32, 64]).to(DEV).eval()
    B, N, S = 2, 300, 40
    xyz1 = synthetic.s_uniform(B, N, 1).permute(0, 2, 1).contiguous().to(DEV)
    xyz2 = xyz1[:, :, :S].contiguous()
    p1 = torch.randn(B, 8, N, device=DEV, requires_grad=True)
    p2 = torch.randn(B, 16, S, device=DEV, requires_grad=True)
    out = fp(xyz1, xyz2, p1, p2)
    g = torch.randn_like(out)
    keep = g.clone()
    out.backward(gradient=g)
    assert torch.equal(g, keep)


# ---- loss block at K = 16 (BASELINE.json configs[4]'s K) ----------------------------------------------------------

def test_loss_block_k16_vs_oracle():
    B, N, K = 3, 4096, 16
    data = synthetic.s_cyl(B, N, K, 21)
    g = torch.Generator().manual_seed(22)
    X_raw = data["normals"] + 0.3 * torch.randn(B, N, 3, generator=g)
    W_raw = torch.randn(B, N, 2 * K, generator=g)
    perm = torch.stack([torch.randperm(K, generator=g) for _ in range(B)])
    col = torch.gather(perm, 1, data["inst"].clamp_min(0)) * 2 + data["bb"]
    W_raw.scatter_add_(2, col[:, :, None], torch.full((B, N, 1), 2.5))
    dev = {k: v.to(DEV) for k, v in data.items()}
    for norm_eig in (False, True):
        out = pipeline.loss_forward(dev["pcs"], X_raw.to(DEV), W_raw.to(DEV), dev["normals"], dev["inst"], dev["bb"],
                                    dev["axes"], dev["centers"], norm_eig=norm_eig)
        ref = orc.loss_block(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"], data["axes"],
                             data["centers"], norm_eig=norm_eig)
        assert int(ref["mask"].sum(1).max()) > 8                  # more than 8 live instances: the K = 16 paths
        assert torch.equal(out["matching_indices"].cpu(), ref["matching_indices"])
        assert torch.equal(out["mask"].cpu(), ref["mask"])
        for k in ("total", "normal", "miou", "bb", "axis", "center"):
            assert rel_err(out[k], ref[k]) <= TOL, (k, norm_eig)
        m = ref["mask"]
        dots = (out["E_AX"].cpu() * ref["E_AX"]).sum(-1).abs()
        assert float((1 - dots[m]).max()) <= TOL
        assert rel_err(out["centers"].cpu()[m], ref["centers"][m]) <= TOL


# ---- in-kernel dropout mask (models/pointnet_extrusion.py:60: F.dropout(p=0.5), always on) ------------------------

def philox_mask(seed, B, N, C=128):
    """Reads the mask the head kernel draws: H = 1, no BN fold, W = 32 selector rows at a time."""
    H = torch.ones(B * N, C, device=DEV)
    cols = []
    for c0 in range(0, C, 32):
        W = torch.zeros(32, C, device=DEV)
        W[torch.arange(32), c0 + torch.arange(32)] = 1.0
        cols.append(ops.head_masked(H, None, None, None, W, None, B, N, seed=seed))
    return torch.cat(cols, dim=1)                                  # (B*N, C)


def test_philox_dropout_mask_statistics_and_determinism():
    B, N, C = 4, 4096, 128
    s1 = torch.tensor([1234567, -42], dtype=torch.long, device=DEV)
    m1 = philox_mask(s1, B, N)
    assert set(m1.unique().tolist()) == {0.0, 2.0}                 # kept values scaled by 1/(1-p)
    n = m1.numel()
    assert abs(float(m1.mean()) - 1.0) <= 4 * 2 * 0.5 / math.sqrt(n) * 1.5      # E[mask] = 1, sd 1/sqrt(n)
    keep = (m1 > 0).float()
    assert float((keep.mean(0) - 0.5).abs().max()) <= 5 * 0.5 / math.sqrt(B * N)   # every channel
    assert float((keep.mean(1) - 0.5).abs().max()) <= 6 * 0.5 / math.sqrt(C)       # every point
    # neighbouring channels / points are uncorrelated
    a, b = keep[:, :-1].reshape(-1) - 0.5, keep[:, 1:].reshape(-1) - 0.5
    assert abs(float((a * b).mean())) * 4 <= 5 / math.sqrt(a.numel())
    a, b = keep[:-1].reshape(-1) - 0.5, keep[1:].reshape(-1) - 0.5
    assert abs(float((a * b).mean())) * 4 <= 5 / math.sqrt(a.numel())
    assert torch.equal(m1, philox_mask(s1.clone(), B, N))          # same words -> same mask
    m2 = philox_mask(torch.tensor([1234568, -42], dtype=torch.long, device=DEV), B, N)
    assert 0.45 <= float((m1 != m2).float().mean()) <= 0.55        # another seed -> an independent mask
    m3 = philox_mask(torch.tensor([1234567, -41], dtype=torch.long, device=DEV), B, N)
    assert 0.45 <= float((m1 != m3).float().mean()) <= 0.55


def test_philox_dropout_backward_regenerates_the_same_mask():
    B, N, C, Nout = 2, 1000, 128, 19
    seed = torch.tensor([99, 7], dtype=torch.long, device=DEV)
    mask = philox_mask(seed, B, N)                                 # (B*N, C)
    g = torch.Generator().manual_seed(0)
    H = torch.randn(B * N, C, generator=g).to(DEV)
    sc, sh = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    W = torch.randn(Nout, C, generator=g).to(DEV)
    bias = torch.randn(Nout, generator=g).to(DEV)
    A_ref = torch.relu(H * sc + sh) * mask
    Y_ref = A_ref.double() @ W.double().t() + bias.double()
    # default precision: the tcgen05 layer kernel draws the mask in its operand transform - the same bits
    Y_tc = ops.head_masked(H, sc, sh, None, W, bias, B, N, seed=seed)
    assert rel_err(Y_tc, Y_ref) <= 1e-5
    Y = ops.head_masked(H, sc, sh, None, W, bias, B, N, seed=seed, precision=_lib.PREC_FP32)
    assert rel_err(Y, Y_ref) <= 1e-5
    # the same through an explicit (B, C, N) mask tensor: bit-identical on the SIMT kernel
    mask_cf = mask.reshape(B, N, C).permute(0, 2, 1).contiguous()
    assert torch.equal(Y, ops.head_masked(H, sc, sh, mask_cf, W, bias, B, N))
    # a pending BatchNorm folded by the head itself, ragged row count, wider C
    for (B2, N2, C2) in ((3, 777, 128), (1, 130, 64), (2, 512, 192)):
        H2 = torch.randn(B2 * N2, C2, generator=g).to(DEV)
        sc2, sh2 = (torch.rand(C2, generator=g) + 0.5).to(DEV), torch.randn(C2, generator=g).to(DEV)
        W2 = torch.randn(Nout, C2, generator=g).to(DEV)
        a = ops.head_masked(H2, sc2, sh2, None, W2, bias, B2, N2, seed=seed)
        b = ops.head_masked(H2, sc2, sh2, None, W2, bias, B2, N2, seed=seed, precision=_lib.PREC_FP32)
        assert rel_err(a, b.double()) <= 1e-5, (B2, N2, C2)
    dOut = torch.randn(B * N, Nout, generator=g).to(DEV)
    dA, A = ops.head_bwd(dOut, None, W, B, N, H, sc, sh, seed=seed)
    dA2, A2 = ops.head_bwd(dOut, mask_cf, W, B, N, H, sc, sh)
    assert torch.equal(dA, dA2) and torch.equal(A, A2)
    assert rel_err(A, A_ref) <= 1e-6
    # the training step's call: upstream rows padded to 16 bytes (the padding is uninitialised memory - NaN here) take
    # the rows-as-warps kernel (lane = four channels); ragged row count (2000 = 62 * 32 + 16); same bits, same FMA order
    d_buf = torch.full((B * N, 20), float("nan"), device=DEV)
    d_buf[:, :Nout] = dOut
    dA3, A3 = ops.head_bwd(d_buf[:, :Nout], None, W, B, N, H, sc, sh, seed=seed)
    assert torch.equal(dA3, dA2) and torch.equal(A3, A2)
    dA4 = ops.head_bwd(d_buf[:, :Nout], None, W, B, N)              # no mask, no re-materialised input
    assert rel_err(dA4, dOut.double() @ W.double()) <= 1e-5


def _small_net(K=4):
    net = backbone(output_sizes=[3, 2 * K])
    net.load_state_dict(orc.init_state_dict((3, 2 * K), 0), strict=True)
    return net.to(DEV)


def test_backbone_default_dropout_follows_the_cuda_generator():
    """No mask tensor: the seed words come from torch's CUDA generator, so torch.manual_seed reproduces a forward and
    consecutive forwards differ; the eval-mode network with dropout differs from the identity-mask run."""
    B, N, K = 2, 1024, 4
    net = _small_net(K).eval()
    x = synthetic.s_cyl(B, N, K, 0)["pcs"].to(DEV)
    starts = [torch.zeros(B, dtype=torch.long, device=DEV), torch.ones(B, dtype=torch.long, device=DEV)]
    assert pipeline.dropout_mask_fn is None
    with torch.no_grad():
        torch.manual_seed(11)
        a = pipeline.backbone_forward(net, x, starts)[1].clone()
        b = pipeline.backbone_forward(net, x, starts)[1].clone()
        torch.manual_seed(11)
        c = pipeline.backbone_forward(net, x, starts)[1].clone()
    assert torch.equal(a, c) and not torch.equal(a, b)
    assert bool(torch.isfinite(a).all())


def test_graph_survives_update_momentum_and_leaves_bn_buffers_alone():
    """ADVICE r1 (low) + VERDICT item 10: constructing the graph does not advance the BatchNorm running statistics;
    a BatchNorm momentum change (update_momentum, train_Point2Cyl_without_sketch.py:357-366) re-captures instead of
    raising, and the replayed step then uses the new momentum."""
    from point2cyl_b200.graph import GraphedForwardLoss
    B, N, K = 2, 1024, 4
    net = _small_net(K).train()
    batch = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, 0).items()}
    before = {k: v.clone() for k, v in net.named_buffers()}
    gfl = GraphedForwardLoss(net, batch)
    for k, v in net.named_buffers():
        assert torch.equal(v, before[k]), k
    gfl()
    torch.cuda.synchronize()
    rm1 = net.bn1.running_mean.clone()
    assert not torch.equal(rm1, before["bn1.running_mean"]) and int(net.bn1.num_batches_tracked) == 1
    for m in net.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = 0.5
    assert gfl.stale()
    out = gfl()                                                    # re-captures
    torch.cuda.synchronize()
    assert gfl.captures == 2 and not gfl.stale() and bool(torch.isfinite(out["losses"]).all())
    assert int(net.bn1.num_batches_tracked) == 2                   # the re-capture itself was not a step
    # running_mean moved by momentum 0.5 towards the batch mean: new = 0.5*old + 0.5*batch
    net2 = _small_net(K).train()
    net2.load_state_dict({k: v for k, v in net.state_dict().items()})
    assert float((net.bn1.running_mean - rm1).abs().max()) > 0


def test_start_ring_reuses_slots_only_after_their_copy():
    from point2cyl_b200.graph import StartRing
    ring = StartRing(8, (1000, 50), torch.device(DEV), depth=2)
    torch.manual_seed(3)
    want = []
    for _ in range(5):
        want.append((torch.randint(0, 1000, (8,)), torch.randint(0, 50, (8,))))
    torch.manual_seed(3)
    for w in want:                                                 # same CPU-generator stream, same order as the reference
        ring.draw()
        assert torch.equal(ring.dev[0].cpu(), w[0]) and torch.equal(ring.dev[1].cpu(), w[1])


def test_geometry_stage_equals_the_fused_calls():
    """pipeline.geometry_forward (+ the search / gather halves of the 3-NN kernel) against the one-kernel calls."""
    B, N = 3, 2048
    net = _small_net(4).eval()
    xyz = synthetic.s_cyl(B, N, 4, 5)["pcs"].to(DEV)
    starts = [torch.tensor([5, 6, 7], device=DEV), torch.tensor([1, 2, 3], device=DEV)]
    geo = pipeline.geometry_forward(net, xyz, starts)
    buf = pipeline.Geometry.empty(net, B, N, torch.device(DEV, torch.cuda.current_device()))
    geo2 = pipeline.geometry_forward(net, xyz, starts, out=buf)
    assert geo2 is buf
    for f in ("fps1", "l1_xyz", "gidx1", "fps2", "l2_xyz", "gidx2", "nn1_idx", "nn1_w", "nn2_idx", "nn2_w", "xyz"):
        assert torch.equal(getattr(geo, f), getattr(buf, f)), f
    feats = torch.randn(B * 512, 20, device=DEV)
    out, idx, w = ops.three_nn_interp(xyz, geo.l1_xyz, feats, want_idx=True)
    assert torch.equal(idx, geo.nn1_idx) and torch.equal(w, geo.nn1_w)
    wide = torch.zeros(B * N, 28, device=DEV)
    ops.three_nn_gather(feats, geo.nn1_idx, geo.nn1_w, 512, out=wide[:, 8:])
    assert torch.equal(wide[:, 8:], out) and bool((wide[:, :8] == 0).all())


def test_pipelined_forward_loss_equals_sequential():
    """graph.PipelinedForwardLoss (geometry of batch i+1 on a second stream under the layers of batch i, SM budget)
    returns what pipeline.forward_loss returns for the same batches in the same order."""
    from point2cyl_b200 import pin_batch
    from point2cyl_b200.graph import PipelinedForwardLoss
    B, N, K = 4, 2048, 4
    batches = [synthetic.s_cyl(B, N, K, 100 + i) for i in range(4)]
    real = pipeline.dropout_mask_fn
    pipeline.dropout_mask_fn = lambda ones, p=0.5: ones
    try:
        net = _small_net(K).train()
        state = {k: v.clone() for k, v in net.state_dict().items()}
        torch.manual_seed(77)
        ref = []
        for b in batches:
            with torch.no_grad():
                o = pipeline.forward_loss(net, {k: v.to(DEV) for k, v in b.items()})
            ref.append({k: o[k].clone() for k in ("losses", "X_raw", "W_raw", "matching_indices")})
        ref_state = {k: v.clone() for k, v in net.state_dict().items()}
        net.load_state_dict(state)
        pipe = PipelinedForwardLoss(net, {k: v.to(DEV) for k, v in batches[0].items()})
        for k, v in net.state_dict().items():                       # capture is not a training step
            assert torch.equal(v, state[k]), k
        torch.manual_seed(77)
        host = [pin_batch(b) for b in batches]
        pipe.prime(host[0])
        for i in range(4):
            out = pipe.step(host[i + 1] if i + 1 < 4 else None)
            torch.cuda.synchronize()
            assert torch.equal(out["matching_indices"], ref[i]["matching_indices"]), i
            # (another grid size = another summation order of the BatchNorm statistics; train-mode BatchNorm on a
            # batch of 4 amplifies that rounding - measured 1.5e-5)
            assert rel_err(out["losses"], ref[i]["losses"]) <= TOL, i
            assert rel_err(out["X_raw"], ref[i]["X_raw"]) <= TOL and rel_err(out["W_raw"], ref[i]["W_raw"]) <= TOL, i
        pipe.join()
        torch.cuda.synchronize()
        for k, v in net.state_dict().items():                       # four real steps of running statistics
            if v.is_floating_point():
                assert rel_err(v, ref_state[k]) <= TOL, k
            else:
                assert torch.equal(v, ref_state[k]), k
        assert ops.set_sm_budget(0) == 0                            # the budget is only set while capturing
    finally:
        pipeline.dropout_mask_fn = real


@pytest.mark.parametrize("partition", ["soft", "green"])
def test_deep_pipelined_forward_loss_equals_sequential(partition):
    """(green: the SMs split hard between the layers and the other two stages with CUDA green contexts.)
    graph.DeepPipelinedForwardLoss (three stages over three slots: geometry of batch i+1 | layers of batch i | loss of
    batch i-1 on a third stream) returns, one call later, what pipeline.forward_loss returns for the same batches in
    the same order - including a run longer than the slot count (slot reuse) and the drained last batch."""
    from point2cyl_b200 import pin_batch
    from point2cyl_b200.graph import DeepPipelinedForwardLoss
    B, N, K = 4, 2048, 4
    nb = 7
    batches = [synthetic.s_cyl(B, N, K, 200 + i) for i in range(nb)]
    real = pipeline.dropout_mask_fn
    pipeline.dropout_mask_fn = lambda ones, p=0.5: ones
    try:
        net = _small_net(K).train()
        state = {k: v.clone() for k, v in net.state_dict().items()}
        torch.manual_seed(78)
        ref = []
        for b in batches:
            with torch.no_grad():
                o = pipeline.forward_loss(net, {k: v.to(DEV) for k, v in b.items()})
            ref.append({k: o[k].clone() for k in ("losses", "X_raw", "W_raw", "matching_indices")})
        ref_state = {k: v.clone() for k, v in net.state_dict().items()}
        net.load_state_dict(state)
        pipe = DeepPipelinedForwardLoss(net, {k: v.to(DEV) for k, v in batches[0].items()}, partition=partition)
        if partition == "green" and pipe.partition != "green":
            pytest.skip("green contexts unavailable (cuda-python driver bindings / driver support)")
        if partition == "green":
            assert pipe.geometry_sms % 8 == 0 and pipe.geometry_sms + pipe.feature_sms == \
                torch.cuda.get_device_properties(0).multi_processor_count
        for k, v in net.state_dict().items():                       # capture is not a training step
            assert torch.equal(v, state[k]), k
        torch.manual_seed(78)
        host = [pin_batch(b) for b in batches]
        pipe.prime(host[0])
        outs = []
        for i in range(nb):
            out = pipe.step(host[i + 1] if i + 1 < nb else None)
            if i == 0:
                assert out is None                                   # the third stage is still empty
            else:
                torch.cuda.synchronize()
                outs.append({k: out[k].clone() for k in ref[0]})
        last = pipe.flush()
        torch.cuda.synchronize()
        outs.append({k: last[k].clone() for k in ref[0]})
        assert len(outs) == nb
        for i, (o, r) in enumerate(zip(outs, ref)):
            assert torch.equal(o["matching_indices"], r["matching_indices"]), i
            assert rel_err(o["losses"], r["losses"]) <= TOL, i
            assert rel_err(o["X_raw"], r["X_raw"]) <= TOL and rel_err(o["W_raw"], r["W_raw"]) <= TOL, i
        pipe.join()
        torch.cuda.synchronize()
        for k, v in net.state_dict().items():
            if v.is_floating_point():
                assert rel_err(v, ref_state[k]) <= TOL, k
            else:
                assert torch.equal(v, ref_state[k]), k
    finally:
        pipeline.dropout_mask_fn = real
