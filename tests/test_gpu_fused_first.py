"""GPU tests of the round-2 fusions of the forward step:
* sa1 without its first layer's activations - closed-form BatchNorm statistics (p2c_group_moments + p2c_sa_xyz_stats)
  and the first conv recomputed in the second layer's operand transform (p2c_sa_xyz_linear),
  models/pointnet_util.py:130-139,200-203;
* the output heads on the tcgen05 layer kernel with the dropout mask drawn in the operand transform
  (p2c_head_masked with P2C_PREC_3XTF32), models/pointnet_extrusion.py:60-65."""
import pytest
import torch

from point2cyl_b200 import _lib, ops, pipeline, synthetic
from point2cyl_b200.dropin.models.pointnet_extrusion import backbone

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _grouping(B, N, S, ns, radius, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(B, N, 3, generator=g).to(DEV)
    start = torch.randint(0, N, (B,), generator=g).to(DEV)
    _, new_xyz = ops.fps(xyz, S, start)
    gidx = ops.ball_query(radius, ns, xyz, new_xyz)
    return xyz, new_xyz, gidx, g


@pytest.mark.parametrize("B,N,S,ns,C0,N1,pool", [(2, 2048, 128, 64, 64, 64, 0), (3, 1000, 50, 32, 64, 128, 32),
                                                 (1, 4096, 512, 64, 64, 64, 0), (2, 512, 33, 16, 128, 128, 0),
                                                 (2, 700, 27, 64, 64, 128, 64)])
def test_sa_xyz_linear_equals_materialised_first_layer(B, N, S, ns, C0, N1, pool):
    xyz, new_xyz, gidx, g = _grouping(B, N, S, ns, 0.2, N + S)
    rows = B * S * ns
    W0 = (torch.randn(C0, 3, generator=g) * 2).to(DEV)
    b0 = torch.randn(C0, generator=g).to(DEV)
    W1 = (torch.randn(N1, C0, generator=g) / C0 ** 0.5).to(DEV)
    b1 = torch.randn(N1, generator=g).to(DEV)
    # closed-form statistics of the first layer == the sums the materialising kernel accumulates
    st_ref = torch.zeros(2 * C0, dtype=torch.float64, device=DEV)
    Y0 = ops.sa_first_layer(xyz, new_xyz, gidx, None, W0, b0, st_ref)
    st = torch.zeros(2 * C0, dtype=torch.float64, device=DEV)
    ops.sa_xyz_stats(ops.group_moments(xyz, new_xyz, gidx), rows, W0, b0, st)
    exact = torch.cat([Y0.double().sum(0), (Y0.double() ** 2).sum(0)])
    assert rel_err(st, exact) <= 1e-6
    assert rel_err(st_ref, exact) <= 1e-6
    mean = st[:C0] / rows
    var = st[C0:] / rows - mean ** 2
    assert float(((var - Y0.double().var(0, unbiased=False)).abs() / Y0.double().var(0, unbiased=False)).max()) <= 1e-5
    # same folded BatchNorm into both
    sc = (torch.rand(C0, generator=g) + 0.5).to(DEV)
    sh = torch.randn(C0, generator=g).to(DEV)
    s1 = torch.zeros(2 * N1, dtype=torch.float64, device=DEV)
    s2 = torch.zeros(2 * N1, dtype=torch.float64, device=DEV)
    assert _lib.load().p2c_linear_path(C0, rows, N1, C0, 0, pool, _lib.PREC_3XTF32, 0) == 1
    ref = ops.linear(Y0, W1, b1, in_scale=sc, in_shift=sh, stats=s1, pool_group=pool, precision=_lib.PREC_3XTF32)
    got = ops.sa_xyz_linear(xyz, new_xyz, gidx, W0, b0, W1, b1, scale0=sc, shift0=sh, stats=s2, pool_group=pool)
    assert got is not None
    # (the BatchNorm is folded into the conv's coefficients: one rounding apart from the materialised operand)
    for g_, r_ in (zip(got, ref) if pool else ((got, ref),)):
        assert rel_err(g_, r_) <= 2e-6
    assert rel_err(s2, s1) <= 2e-6
    # and against float64 of the definition
    A = torch.relu(Y0.double() * sc.double() + sh.double())
    Yd = A @ W1.double().t() + b1.double()
    assert rel_err(got[0] if pool else got, Yd) <= 1e-5


def _net(K=4):
    torch.manual_seed(5)
    return backbone(output_sizes=[3, 2 * K]).to(DEV)


@pytest.mark.parametrize("train", [True, False])
def test_backbone_with_and_without_materialised_sa1(train):
    """pipeline.xyz_first_enabled switches sa1 between the two paths: same outputs, same running statistics."""
    B, N, K = 4, 2048, 4
    batch = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, 11).items()}
    starts = (torch.zeros(B, dtype=torch.long, device=DEV), torch.zeros(B, dtype=torch.long, device=DEV))
    real_mask, real_flag = pipeline.dropout_mask_fn, pipeline.xyz_first_enabled
    pipeline.dropout_mask_fn = lambda ones, p=0.5: ones
    try:
        outs, states = [], []
        for flag in (False, True):
            net = _net(K).train(train)
            pipeline.xyz_first_enabled = flag
            with torch.no_grad():
                c0 = _lib.launch_count
                o = pipeline.forward_loss(net, batch, fps_start=starts)
                n_launch = _lib.launch_count - c0
            outs.append((o, n_launch))
            states.append({k: v.clone() for k, v in net.state_dict().items()})
        (a, la), (b, lb) = outs
        for k in ("X_raw", "W_raw"):
            assert rel_err(b[k], a[k]) <= 1e-4, k   # train-mode BatchNorm on 4 clouds amplifies the 1e-7 difference of the statistics (measured 6.8e-5)
        # (the axis term is left out: see test_backbone_with_and_without_fp3_concat)
        assert rel_err(b["losses"][[1, 2, 3, 5]], a["losses"][[1, 2, 3, 5]]) <= 1e-4
        assert torch.equal(a["matching_indices"], b["matching_indices"])
        for k, v in states[0].items():
            if v.is_floating_point():
                assert rel_err(states[1][k], v) <= 1e-4, k
            else:
                assert torch.equal(states[1][k], v), k
    finally:
        pipeline.dropout_mask_fn, pipeline.xyz_first_enabled = real_mask, real_flag


def test_head_on_tensor_cores_matches_simt_head_inside_the_backbone():
    """fp32 (SIMT head) vs 3xtf32 (tcgen05 head, mask drawn in the operand transform) with the same dropout seed."""
    B, N, C, Nout = 3, 1500, 128, 19
    g = torch.Generator().manual_seed(1)
    H = torch.randn(B * N, C, generator=g).to(DEV)
    W = (torch.randn(Nout, C, generator=g) / C ** 0.5).to(DEV)
    bias = torch.randn(Nout, generator=g).to(DEV)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    seed = torch.tensor([31337, 4242], dtype=torch.long, device=DEV)
    outs = []
    for prec in (_lib.PREC_FP32, _lib.PREC_3XTF32):
        stats = torch.cat([H.double().sum(0), (H.double() ** 2).sum(0)])
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        bn = ops.PendingBN(stats, B * N, gamma, beta, 1e-5, 0.1, True, rm, rv, save=True)
        outs.append((ops.head_masked(H, None, None, None, W, bias, B, N, seed=seed, bn=bn, precision=prec),
                     bn.scale.clone(), bn.mean.clone(), rm, rv))
    assert rel_err(outs[1][0], outs[0][0]) <= 1e-5
    for i in (1, 2, 3, 4):            # the folded BatchNorm was published and the running statistics updated once
        assert torch.equal(outs[0][i], outs[1][i])


# ---- fp3 through the linearity of its first conv (models/pointnet_util.py:298-299, 312, 317) -------------------------

@pytest.mark.parametrize("M,N,K", [(32, 256, 1024), (5, 19, 100), (1, 64, 3072), (64, 130, 257)])
def test_linear_small(M, N, K):
    g = torch.Generator().manual_seed(M + N)
    X = torch.randn(M, K, generator=g).to(DEV)
    Wwide = torch.randn(N, K + 40, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    got = ops.linear_small(X, Wwide[:, 40:], bias)                 # a column slice of a wider weight matrix
    assert rel_err(got, X.double() @ Wwide[:, 40:].double().t() + bias.double()) <= 2e-6


@pytest.mark.parametrize("B,N,D1,D2,C0,affine", [(4, 128, 256, 1024, 256, False), (3, 96, 256, 300, 128, True),
                                                 (2, 160, 512, 64, 256, False)])
def test_linear_group_bias_equals_concat_layer(B, N, D1, D2, C0, affine):
    g = torch.Generator().manual_seed(D1 + D2)
    M = B * N
    X = torch.randn(M, D1, generator=g).to(DEV)
    V = torch.randn(B, D2, generator=g).to(DEV)
    W = (torch.randn(C0, D1 + D2, generator=g) / (D1 + D2) ** 0.5).to(DEV)
    bias = torch.randn(C0, generator=g).to(DEV)
    sc = (torch.rand(D1, generator=g) + 0.5).to(DEV) if affine else None
    sh = torch.randn(D1, generator=g).to(DEV) if affine else None
    per = ops.linear_small(V, W[:, D1:], bias)
    ws = ops.split_tf32_multi([W[:, :D1]])[0]
    stats = torch.zeros(2 * C0, dtype=torch.float64, device=DEV)
    got = ops.linear_group_bias(X, ws, per, N, C0, D1, in_scale=sc, in_shift=sh, stats=stats)
    A = X.double() if not affine else torch.relu(X.double() * sc.double() + sh.double())
    full = torch.cat([A, V.double().repeat_interleave(N, 0)], 1)
    ref = full @ W.double().t() + bias.double()
    assert rel_err(got, ref) <= 1e-5
    assert rel_err(stats, torch.cat([ref.sum(0), (ref ** 2).sum(0)])) <= 1e-5


@pytest.mark.parametrize("train", [True, False])
def test_backbone_with_and_without_fp3_concat(train):
    B, N, K = 4, 2048, 4
    batch = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, 12).items()}
    starts = (torch.zeros(B, dtype=torch.long, device=DEV), torch.zeros(B, dtype=torch.long, device=DEV))
    real_mask, real_flag = pipeline.dropout_mask_fn, pipeline.group_bias_enabled
    pipeline.dropout_mask_fn = lambda ones, p=0.5: ones
    try:
        outs, states = [], []
        for flag in (False, True, True):          # the second enabled pass splits the slice with the other weights
            if flag is False or not outs[-1][1]:
                net = _net(K).train(train)
            pipeline.group_bias_enabled = flag
            with torch.no_grad():
                o = pipeline.forward_loss(net, batch, fps_start=starts)
            outs.append((o, flag))
            states.append({k: v.clone() for k, v in net.state_dict().items()})
        assert net._p2c_split_slices, "the linearity path did not run"
        a, b = outs[0][0], outs[1][0]
        for k in ("X_raw", "W_raw"):
            assert rel_err(b[k], a[k]) <= 1e-4, k
        # (not the axis term: a random-init network's fitted axes sit at near-degenerate eigenvalues, where a 1e-6
        # change of the inputs picks another eigenvector - measured 0.62 vs 0.68 in eval mode)
        assert rel_err(b["losses"][[1, 2, 3, 5]], a["losses"][[1, 2, 3, 5]]) <= 1e-4
        assert torch.equal(a["matching_indices"], b["matching_indices"])
        for k, v in states[0].items():
            if v.is_floating_point():
                assert rel_err(states[1][k], v) <= 1e-4, k
        if not train:                            # eval mode: no state changes, the second enabled pass repeats the first
            assert torch.equal(outs[2][0]["X_raw"], b["X_raw"])
    finally:
        pipeline.dropout_mask_fn, pipeline.group_bias_enabled = real_mask, real_flag


# ---- a whole feature-less SA level as ONE kernel in eval mode (p2c_sa_stack_fused, csrc/sa_stack_tc.cu) -------------


def _sa_module(C2, ns, S, radius, seed):
    from point2cyl_b200.dropin.models.pointnet_util import PointNetSetAbstraction
    torch.manual_seed(seed)
    sa = PointNetSetAbstraction(npoint=S, radius=radius, nsample=ns, in_channel=3, mlp=[64, 64, C2], group_all=False)
    g = torch.Generator().manual_seed(seed + 1)
    for bn in sa.mlp_bns:                       # non-trivial running statistics, some NEGATIVE scales (the min branch)
        C_ = bn.weight.shape[0]
        bn.weight.data = (torch.rand(C_, generator=g) + 0.5) * torch.where(torch.rand(C_, generator=g) < 0.25, -1.0, 1.0)
        bn.bias.data = torch.randn(C_, generator=g) * 0.3
        bn.running_mean.data = torch.randn(C_, generator=g) * 0.2
        bn.running_var.data = torch.rand(C_, generator=g) + 0.3
    return sa.to(DEV).eval()


@pytest.mark.parametrize("B,N,S,ns,C2", [(2, 2048, 128, 64, 128), (3, 1000, 50, 32, 128), (1, 4096, 512, 64, 128),
                                         (2, 700, 27, 64, 96), (2, 900, 9, 128, 128), (1, 300, 1, 32, 64)])
def test_sa_stack_fused_equals_per_layer_path_and_fp64(B, N, S, ns, C2):
    """The one-kernel level against (a) the per-layer kernels (p2c_sa_xyz_linear + p2c_linear + p2c_pool_bn_relu) and
    (b) a float64 restatement of models/pointnet_util.py:130-139, 200-205 in eval mode, partial last tile and partial
    channel count included."""
    sa = _sa_module(C2, ns, S, 0.25, seed=N + S)
    xyz, new_xyz, gidx, g = _grouping(B, N, S, ns, 0.25, N + S)
    with torch.no_grad():
        assert pipeline.stack_fused_enabled
        _, fused = pipeline.set_abstraction(sa, xyz, None, None, geo=(None, new_xyz, gidx))
        pipeline.stack_fused_enabled = False
        try:
            _, layered = pipeline.set_abstraction(sa, xyz, None, None, geo=(None, new_xyz, gidx))
        finally:
            pipeline.stack_fused_enabled = True
    assert fused.shape == (B * S, C2) and layered.shape == fused.shape
    # float64 restatement
    idx = gidx.cpu()
    x = xyz.cpu().double()
    grouped = torch.stack([x[b][idx[b]] for b in range(B)]) - new_xyz.cpu().double()[:, :, None, :]   # (B,S,ns,3)
    h = grouped.reshape(B * S * ns, 3)
    for conv, bn in zip(sa.mlp_convs, sa.mlp_bns):
        W = conv.weight.detach().cpu().double().reshape(conv.weight.shape[0], -1)
        h = h @ W.t() + conv.bias.detach().cpu().double()
        h = (h - bn.running_mean.cpu().double()) / torch.sqrt(bn.running_var.cpu().double() + bn.eps) \
            * bn.weight.detach().cpu().double() + bn.bias.detach().cpu().double()
        h = torch.relu(h)
    exact = h.reshape(B * S, ns, C2).max(dim=1)[0]
    assert rel_err(fused, exact) <= 1e-5
    assert rel_err(layered, exact) <= 1e-5
    assert rel_err(fused, layered) <= 1e-5


def test_eval_backbone_uses_the_fused_level_and_matches_per_layer():
    """Whole backbone in eval mode at a config-2-like shape: with and without the fused sa1 kernel."""
    B, N, K = 2, 8192, 8
    data = synthetic.s_cyl(B, N, K, seed=11)
    torch.manual_seed(3)
    net = backbone(output_sizes=[3, 2 * K]).to(DEV).eval()
    pcs = data["pcs"].to(DEV)
    start = [torch.randint(0, N, (B,)).to(DEV), torch.randint(0, 512, (B,)).to(DEV)]
    mask = ((torch.rand(B, 128, N) > 0.5).float() * 2.0).to(DEV)      # the head's dropout is always on (:60): fix it
    real = pipeline.dropout_mask_fn
    pipeline.dropout_mask_fn = lambda ones, p=0.5: mask
    try:
        _lib.profile_start()
        with torch.no_grad():
            a = pipeline.backbone_forward(net, pcs, start)
        names = [n for n, _, _ in _lib.profile_stop()]
        assert "p2c_sa_stack_fused" in names and "p2c_sa_xyz_linear" not in names
        pipeline.stack_fused_enabled = False
        with torch.no_grad():
            b = pipeline.backbone_forward(net, pcs, start)
    finally:
        pipeline.stack_fused_enabled = True
        pipeline.dropout_mask_fn = real
    for u, v in zip(a, b):
        assert rel_err(u, v) <= 1e-5
