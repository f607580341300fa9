"""GPU parity tests of the backward kernels: gradients through the C-ABI against the
reference's own autograd (goldens) and against autograd through the CPU oracle.  Gradient tolerance: 1e-4 of the
largest reference entry per tensor unless stated."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import p2c_oracle as orc
from point2cyl_b200 import ops, pipeline, synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda"
LOSS = ["loss_b2_n1024_k4.npz", "loss_b3_n2048_k8_normeig.npz"]


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", LOSS)
def test_fused_loss_backward_golden(golden_dir, name):
    g = load(golden_dir, name)
    B, N, K, seed, norm_eig = (int(v) for v in g["meta"])
    data = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed).items()}
    X_raw = torch.from_numpy(g["X_raw"]).to(DEV).requires_grad_(True)
    W_raw = torch.from_numpy(g["W_raw"]).to(DEV).requires_grad_(True)
    out = pipeline.loss_forward(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"], data["axes"],
                                data["centers"], norm_eig=bool(norm_eig))
    assert rel_err(out["total"], float(g["total"]) + float(g["bb"]) + float(g["axis"]) + float(g["center"])) <= TOL
    out["total"].backward()
    assert rel_err(X_raw.grad, g["dX_raw"]) <= TOL
    assert rel_err(W_raw.grad, g["dW_raw"]) <= TOL


@pytest.mark.parametrize("term", range(5))
def test_fused_loss_backward_per_term(term):
    """Each loss term alone (one multiplier = 1) against autograd through the oracle, K = 8 with empty slots."""
    B, N, K = 3, 1500, 8
    data = synthetic.s_cyl(B, N, K, seed=11)
    g = torch.Generator().manual_seed(2)
    X_raw = (data["normals"] + 0.3 * torch.randn(B, N, 3, generator=g)) * 1.7
    W_raw = torch.randn(B, N, 2 * K, generator=g)
    col = data["inst"] * 2 + data["bb"]
    W_raw.scatter_add_(2, col[:, :, None], torch.full((B, N, 1), 2.0))
    weights = [0.0] * 5
    weights[term] = 1.0
    Xc, Wc = X_raw.clone().requires_grad_(True), W_raw.clone().requires_grad_(True)
    ref = orc.loss_block(data["pcs"], Xc, Wc, data["normals"], data["inst"], data["bb"], data["axes"],
                         data["centers"], weights=tuple(weights))
    ref["total"].backward()
    Xd, Wd = X_raw.to(DEV).requires_grad_(True), W_raw.to(DEV).requires_grad_(True)
    d = {k: v.to(DEV) for k, v in data.items()}
    out = pipeline.loss_forward(d["pcs"], Xd, Wd, d["normals"], d["inst"], d["bb"], d["axes"], d["centers"],
                                weights=tuple(weights))
    out["total"].backward()
    assert torch.equal(out["matching_indices"].cpu(), ref["matching_indices"])
    assert rel_err(out["total"], ref["total"]) <= TOL
    for got, want in ((Xd.grad, Xc.grad), (Wd.grad, Wc.grad)):
        if float(want.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0
        else:
            assert rel_err(got, want) <= TOL


@pytest.mark.parametrize("name", LOSS)
def test_function_level_backward_golden(golden_dir, name):
    """The training script's own sequence (train_Point2Cyl_without_sketch.py:246-353) on the drop-in functions with
    torch autograd for the inline glue: same gradients as the reference."""
    from point2cyl_b200.dropin import data_utils as du
    from point2cyl_b200.dropin import losses as ls
    g = load(golden_dir, name)
    B, N, K, seed, norm_eig = (int(v) for v in g["meta"])
    data = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed).items()}
    X_raw = torch.from_numpy(g["X_raw"]).to(DEV).requires_grad_(True)
    W_raw = torch.from_numpy(g["W_raw"]).to(DEV).requires_grad_(True)
    X = F.normalize(X_raw, p=2, dim=2, eps=1e-12)
    W_2K = torch.softmax(W_raw, dim=2)
    W_barrel, W_barrel_bb = W_2K[:, :, ::2], W_raw[:, :, ::2]
    W_base, W_base_bb = W_2K[:, :, 1::2], W_raw[:, :, 1::2]
    W = W_barrel + W_base
    total, l_n, l_seg, match, mask = ls.compute_all_losses(data["pcs"], W, data["inst"], X, data["normals"], 1.0, 1.0,
                                                           return_match_indices=True)
    l_bb = _inline_bb(W, W_barrel_bb, W_base_bb, data["bb"], match, mask, K)
    mask_gt = ls.get_mask_gt(data["inst"], K)
    gi = match.unsqueeze(1).expand(B, N, K)
    E_AX = du.estimate_extrusion_axis(X, torch.gather(W_barrel, 2, gi), torch.gather(W_base, 2, gi), data["bb"],
                                      data["inst"], normalize=bool(norm_eig))
    ext = ls.compute_normal_loss(E_AX, data["axes"], angle_diff=False, collapse=False)
    l_ax = torch.mean(ls.reduce_mean_masked_instance(ext, mask_gt))
    centers = du.estimate_extrusion_centers(torch.gather(W, 2, gi), data["pcs"])
    l_c = torch.mean(ls.reduce_mean_masked_instance(torch.square(centers - data["centers"]).sum(dim=-1), mask_gt))
    (total + l_bb + l_ax + l_c).backward()
    assert rel_err(l_ax, g["axis"]) <= TOL and rel_err(l_c, g["center"]) <= TOL and rel_err(l_bb, g["bb"]) <= TOL
    assert rel_err(X_raw.grad, g["dX_raw"]) <= TOL
    assert rel_err(W_raw.grad, g["dW_raw"]) <= TOL


def _inline_bb(W, W_barrel_bb, W_base_bb, gt_bb, match, mask, K):
    """train_Point2Cyl_without_sketch.py:286-307 as torch ops (what the unmodified script runs inline)."""
    B, N, _ = W.shape
    Wr = torch.gather(W, 2, match.unsqueeze(1).expand(B, N, K))
    Wr = torch.where(mask.float().unsqueeze(1).expand(B, N, K) == 1, Wr, torch.zeros_like(Wr))
    Wr = torch.softmax(Wr, dim=-1)
    W_sorted, label = torch.sort(Wr, dim=-1)
    seg = torch.cat((torch.gather(W_barrel_bb, 2, label).unsqueeze(-1), torch.gather(W_base_bb, 2, label).unsqueeze(-1)),
                    dim=-1)
    ce = F.cross_entropy(seg.contiguous().view(B * N * K, -1), gt_bb.unsqueeze(-1).repeat(1, 1, K).view(B * N * K),
                         reduction="none").view(B, N, K)
    return torch.mean(torch.mean(torch.sum(ce * W_sorted, dim=-1), dim=-1))


def test_eig3x3_backward_matches_eigh():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(64, 3, 3, generator=g)
    M = (A @ A.transpose(1, 2)).double().requires_grad_(True)
    gv = torch.randn(64, 3, generator=g)
    e, v = torch.linalg.eigh(M)
    v0 = v[:, :, 0]
    from point2cyl_b200 import autograd as ag
    Md = M.detach().float().to(DEV).requires_grad_(True)
    vec, _ = ag.eig3x3_smallest(Md)
    sign = torch.sign((vec.detach().cpu().double() * v0.detach()).sum(-1, keepdim=True))
    (v0 * sign * gv.double()).sum().backward()
    (vec * gv.to(DEV)).sum().backward()
    ref = 0.5 * (M.grad + M.grad.transpose(1, 2))
    assert rel_err(Md.grad, ref) <= 1e-4


# ---- backbone backward -------------------------------------------------------------------------------------------
#
# Conditioning (measured with the CPU oracle, tests/tools/grad_sensitivity.py): with TRAIN-mode BatchNorm a relative
# perturbation of 1e-6 of the weights changes the reference's own parameter gradients by 1-2 % in L2 (every layer
# upstream of the heads); with BatchNorm on running statistics the same perturbation moves them by <= 5e-4 (isolated
# ReLU / arg-max flips, median 0).  No implementation with a different summation order can therefore match the
# train-mode gradients of the reference tighter than a few per cent, so the end-to-end bars are:
#   * BatchNorm on running statistics: relative L2 error <= 5e-3 per parameter, median entry error <= 2e-4 of max;
#   * train-mode BatchNorm: relative L2 error <= 2e-1 per parameter (2048-entry samples; measured worst 8.8e-2),
#     heads (fc2.*) <= 1e-2 (measured 3.3e-3), loss <= 1e-3;
# and every backward kernel is checked on its own against torch autograd to 1e-4 (tests further down).

from tests.test_oracle_golden import grad_errors, live_keys  # noqa: E402


def _train_setup(g, monkeypatch, training=True):
    from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
    B, N, K, seed = (int(v) for v in g["meta"])
    data = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed).items()}
    net = backbone(output_sizes=[3, 2 * K])
    net.load_state_dict(orc.init_state_dict((3, 2 * K), seed=seed), strict=True)
    net = net.to(DEV).train(training)
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0
    mask = mask.to(DEV)
    monkeypatch.setattr(pipeline, "dropout_mask_fn", lambda x, p=0.5, **kw: mask)
    starts = (torch.from_numpy(g["s1"]).to(DEV), torch.from_numpy(g["s2"]).to(DEV))
    return net, data, starts, K


def _assert_param_grads(named_grads, g, training):
    named = dict(named_grads)
    worst = []
    for k in live_keys(g, training):
        l2, med, nerr = grad_errors(named[k], g, k)
        if training:
            bar = 1e-2 if k.startswith("fc2") else 2e-1
            ok = l2 <= bar and nerr <= bar
        else:
            ok = l2 <= 5e-3 and med <= 2e-4 and nerr <= 5e-3   # measured worst: 2.1e-3, 6e-5, 2.4e-4
        if not ok:
            worst.append((k, l2, med, nerr))
    assert not worst, worst


@pytest.mark.parametrize("name,training", [("train_bneval_b2_n1024_k4.npz", False), ("train_b2_n1024_k4.npz", True)])
def test_backbone_backward_autograd_golden(golden_dir, monkeypatch, name, training):
    """Drop-in module + fused loss + torch's loss.backward(): parameter gradients of one training step against the
    reference's (train_Point2Cyl_without_sketch.py:244-367)."""
    g = load(golden_dir, name)
    net, data, starts, K = _train_setup(g, monkeypatch, training)
    X_raw, W_raw = net(data["pcs"], fps_start=starts)
    assert X_raw.requires_grad and W_raw.requires_grad
    ftol = 1e-3 if training else TOL          # train-mode forward conditioning: see tests/test_gpu_parity.py TRAIN_TOL
    assert rel_err(X_raw, g["X_raw"]) <= ftol and rel_err(W_raw, g["W_raw"]) <= ftol
    out = pipeline.loss_forward(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"], data["axes"],
                                data["centers"])
    assert rel_err(out["total"], g["loss"]) <= ftol
    out["total"].backward()
    _assert_param_grads([(k, p.grad) for k, p in net.named_parameters()], g, training)


def test_trainer_forward_backward_golden(golden_dir, monkeypatch):
    """The flat-buffer Trainer (no torch autograd) produces the same gradients, and its Adam step equals
    torch.optim.Adam on them."""
    from point2cyl_b200.train import Trainer
    g = load(golden_dir, "train_bneval_b2_n1024_k4.npz")
    net, data, starts, K = _train_setup(g, monkeypatch, training=False)
    tr = Trainer(net, lr=1e-3)
    assert all(p.data_ptr() >= tr.flat_param.data_ptr() for p in net.parameters())
    out = tr.forward_backward(data, fps_start=starts)
    assert rel_err(out["total"], g["loss"]) <= TOL
    _assert_param_grads([(k, p.grad) for k, p in net.named_parameters()], g, False)
    # Adam: two steps against torch.optim.Adam fed the same gradients
    p0 = tr.flat_param.clone()
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    for _ in range(2):
        grads = tr.flat_grad.clone()
        tr.step_count += 1
        ops.adam_step(tr.flat_param, tr.flat_grad, tr.exp_avg, tr.exp_avg_sq, tr.lr, tr.step_count)
        ref_p.grad = grads
        opt.step()
    assert float((tr.flat_param - ref_p.detach()).abs().max()) <= 2e-7
    assert float((tr.flat_param - p0).abs().max()) > 1e-4


def test_training_reduces_loss(monkeypatch):
    """Ten Trainer steps on one fixed batch (train-mode BatchNorm, live dropout): the total loss goes down."""
    from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
    from point2cyl_b200.train import Trainer
    B, N, K = 4, 2048, 4
    data = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed=1).items()}
    torch.manual_seed(0)
    net = backbone(output_sizes=[3, 2 * K]).to(DEV).train()
    tr = Trainer(net, lr=1e-3)
    starts = (torch.zeros(B, dtype=torch.long, device=DEV), torch.zeros(B, dtype=torch.long, device=DEV))
    losses = [float(tr.step(data, fps_start=starts)["total"]) for _ in range(10)]
    assert all(np.isfinite(losses))
    assert min(losses[-3:]) < losses[0] - 0.05, losses


def test_interp_and_sa_first_backward_vs_autograd():
    """Scatter-add kernels against torch autograd of their forward definitions."""
    g = torch.Generator().manual_seed(5)
    B, N, S, D = 2, 700, 96, 64
    xyz1, xyz2 = torch.rand(B, N, 3, generator=g), torch.rand(B, S, 3, generator=g)
    feats2 = torch.randn(B * S, D, generator=g)
    out, idx, w = ops.three_nn_interp(xyz1.to(DEV), xyz2.to(DEV), feats2.to(DEV), want_idx=True)
    dI = torch.randn(B * N, D, generator=g)
    dF = ops.three_nn_interp_bwd(dI.to(DEV), idx, w, B, N, S)
    f = feats2.double().reshape(B, S, D).requires_grad_(True)
    gathered = torch.gather(f.unsqueeze(1).expand(B, N, S, D), 2, idx.cpu().unsqueeze(-1).expand(B, N, 3, D))
    ((gathered * w.cpu().double().unsqueeze(-1)).sum(2).reshape(B * N, D) * dI.double()).sum().backward()
    assert rel_err(dF, f.grad.reshape(B * S, D)) <= 1e-5
    dF1 = ops.three_nn_interp_bwd(dI.to(DEV), None, None, B, N, 1)            # broadcast case
    assert rel_err(dF1, dI.double().reshape(B, N, D).sum(1)) <= 1e-5
    # fused gather + first SA conv
    ns, C, Dq = 16, 128, 128
    gidx = torch.randint(0, N, (B, S, ns), generator=g)
    Wx = torch.randn(C, 3 + Dq, generator=g)
    dY = torch.randn(B * S * ns, C, generator=g)
    dQf = torch.zeros(B * N, C, device=DEV)
    dW = torch.zeros(C, 3 + Dq, device=DEV)
    dbias = torch.zeros(C, device=DEV)
    ops.sa_first_bwd(dY.to(DEV), xyz1.to(DEV), xyz2.to(DEV), gidx.to(DEV), dQf, dW, dbias)
    rel = (torch.gather(xyz1.unsqueeze(1).expand(B, S, N, 3), 2, gidx.unsqueeze(-1).expand(B, S, ns, 3))
           - xyz2.unsqueeze(2)).reshape(-1, 3).double()
    assert rel_err(dW[:, :3], dY.double().T @ rel) <= 1e-5
    assert float(dW[:, 3:].abs().max()) == 0.0
    assert rel_err(dbias, dY.double().sum(0)) <= 1e-5
    ref_q = torch.zeros(B * N, C, dtype=torch.float64)
    rows = (torch.arange(B).view(B, 1, 1) * N + gidx).reshape(-1)
    ref_q.index_add_(0, rows, dY.double())
    assert rel_err(dQf, ref_q) <= 1e-5


def test_head_backward_vs_autograd():
    g = torch.Generator().manual_seed(6)
    B, N, C, Nout = 2, 600, 128, 19
    H = torch.randn(B * N, C, generator=g)
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    mask = (torch.rand(B, C, N, generator=g) > 0.5).float() * 2
    W = torch.randn(Nout, C, generator=g)
    dOut = torch.randn(B * N, Nout, generator=g)
    Hd = H.double().requires_grad_(True)
    Wd = W.double().requires_grad_(True)
    A = torch.relu(Hd * sc.double() + sh.double()) * mask.double().permute(0, 2, 1).reshape(B * N, C)
    ((A @ Wd.T) * dOut.double()).sum().backward()
    dA = ops.head_bwd(dOut.to(DEV), mask.to(DEV), W.to(DEV), B, N)
    dA2, A_h = ops.head_bwd(dOut.to(DEV), mask.to(DEV), W.to(DEV), B, N, H.to(DEV), sc.to(DEV), sh.to(DEV))
    assert torch.equal(dA, dA2) and rel_err(A_h, A) <= 1e-6
    # dA is the gradient w.r.t. relu(bn(H)) (mask applied); gate it like the BN/ReLU stage does
    gate = ((H * sc + sh) > 0).double() * sc.double()
    assert rel_err(dA.cpu().double() * gate, Hd.grad) <= 1e-5
    dW = torch.zeros(Nout, C, device=DEV)
    db = torch.zeros(Nout, device=DEV)
    ops.wgrad(dOut.to(DEV), H.to(DEV), C, dW, db, sc.to(DEV), sh.to(DEV), mask.to(DEV))
    assert rel_err(dW, Wd.grad) <= 1e-5 and rel_err(db, dOut.double().sum(0)) <= 1e-5


@pytest.mark.parametrize("M,N,K", [(5000, 64, 64), (3000, 128, 131), (1024, 256, 259), (700, 19, 128),
                                   (40000, 128, 128), (33001, 64, 64), (20000, 256, 131), (4096, 512, 1024),
                                   (30001, 128, 64), (9000, 32, 64), (7003, 19, 64)])   # K <= 64: narrow / folded stages
@pytest.mark.parametrize("tc", [False, True])
def test_wgrad_kernel(M, N, K, tc):
    """dW = dY^T relu(bn(X)), db = column sums: fp32 SIMT kernel and the tcgen05 split-over-rows kernel (3xTF32)."""
    from point2cyl_b200 import _lib
    g = torch.Generator().manual_seed(M)
    ldn, ld = ops.pad4(N), ops.pad4(K)
    dY = torch.randn(M, ldn, generator=g)[:, :N]
    X = torch.randn(M, ld, generator=g)
    sc, sh = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.1
    A = torch.relu(X[:, :K] * sc + sh)
    dW = torch.zeros(N, K, device=DEV)
    db = torch.zeros(N, device=DEV)
    dYd, Xd = dY.to(DEV), X.to(DEV)                       # .to keeps the padded row stride of the slice
    if dYd.stride(0) != ldn:
        dYd = torch.zeros(M, ldn, device=DEV)[:, :N].copy_(dY)
    on_tc = ops.wgrad_on_tensor_cores(dYd, Xd, K)
    assert on_tc == (M >= 1024)
    ops.wgrad(dYd, Xd, K, dW, db, sc.to(DEV), sh.to(DEV), precision=_lib.PREC_3XTF32 if tc else _lib.PREC_FP32)
    assert rel_err(dW, dY.double().T @ A.double()) <= 1e-5
    assert rel_err(db, dY.double().sum(0)) <= 1e-5
    # no-affine operand and accumulation into a strided dW view (the Qf half of a fused first layer)
    dWs = torch.ones(N, K + 3, device=DEV)
    ops.wgrad(dYd, Xd, K, dWs[:, 3:], None, precision=_lib.PREC_3XTF32 if tc else _lib.PREC_FP32)
    assert rel_err(dWs[:, 3:] - 1.0, dY.double().T @ X[:, :K].double()) <= 1e-5
    assert float((dWs[:, :3] - 1.0).abs().max()) == 0.0


def test_bn_relu_backward_kernels_vs_autograd():
    """reduce / coef / apply chain against torch autograd of relu(batch_norm(Y)) in float64, plain and pooled."""
    g = torch.Generator().manual_seed(3)
    M, C, G = 4096, 128, 64
    Y = torch.randn(M, C, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g) * 0.2
    dA = torch.randn(M, C, generator=g)
    Yd = Y.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    out = torch.relu(torch.nn.functional.batch_norm(Yd, None, None, gd, bd, True, 0.1, 1e-5))
    (out * dA.double()).sum().backward()
    mean, var = Y.double().mean(0), Y.double().var(0, unbiased=False)
    invstd = 1 / torch.sqrt(var + 1e-5)
    scale, shift = (gamma.double() * invstd).float(), (beta.double() - mean * gamma.double() * invstd).float()
    dev = lambda t: t.float().to(DEV)
    sums = ops.bn_bwd_reduce(dev(dA), dev(Y), dev(scale), dev(shift))
    dgam, dbet = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    coef = ops.bn_bwd_coef(sums, M, dev(gamma), dev(mean), dev(invstd), True, dgam, dbet)
    dY = ops.bn_bwd_apply(dev(dA), dev(Y), dev(scale), dev(shift), coef)
    assert rel_err(dY, Yd.grad) <= 1e-4 and rel_err(dgam, gd.grad) <= 1e-4 and rel_err(dbet, bd.grad) <= 1e-4
    # pooled: max over groups of G rows
    Yd2 = Y.double().requires_grad_(True)
    gd2, bd2 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    act = torch.relu(torch.nn.functional.batch_norm(Yd2, None, None, gd2, bd2, True, 0.1, 1e-5))
    pooled = act.reshape(M // G, G, C).max(dim=1)[0]
    dO = torch.randn(M // G, C, generator=g)
    (pooled * dO.double()).sum().backward()
    Y3 = Y.reshape(M // G, G, C)
    Ymax, Ymin = Y3.max(1)[0], Y3.min(1)[0]
    sums = ops.pool_bwd_reduce(dev(dO), dev(Ymax), dev(Ymin), dev(scale), dev(shift))
    dgam.zero_(), dbet.zero_()
    coef = ops.bn_bwd_coef(sums, M, dev(gamma), dev(mean), dev(invstd), True, dgam, dbet)
    dY = ops.pool_bwd_apply(dev(dO), dev(Ymax), dev(Ymin), dev(Y), dev(scale), dev(shift), coef, G)
    assert rel_err(dY, Yd2.grad) <= 1e-4 and rel_err(dgam, gd2.grad) <= 1e-4 and rel_err(dbet, bd2.grad) <= 1e-4


def test_graphed_trainer_matches_eager(monkeypatch):
    """CUDA-graph replay of forward+loss+backward gives the gradients of the eager Trainer (same dropout mask, same FPS
    starts), keeps BatchNorm buffers untouched by the capture, and trains."""
    from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
    from point2cyl_b200.train import GraphedTrainer, Trainer
    B, N, K = 4, 2048, 4
    data = {k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed=2).items()}
    mask = ((torch.rand(B, 128, N, generator=torch.Generator().manual_seed(1)) > 0.5).float() * 2.0).to(DEV)
    monkeypatch.setattr(pipeline, "dropout_mask_fn", lambda x, p=0.5, **kw: mask)
    starts = (torch.arange(B, device=DEV), torch.arange(B, device=DEV) + 3)
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        nets.append(backbone(output_sizes=[3, 2 * K]).to(DEV).eval())   # running-statistics BatchNorm: well conditioned
    eager = Trainer(nets[0], lr=1e-3)
    graphed = GraphedTrainer(nets[1], data, lr=1e-3)
    for (k, a), (_, b) in zip(nets[0].named_buffers(), nets[1].named_buffers()):
        assert torch.equal(a, b), k                        # capture + warm-up left the running statistics alone
    o1 = eager.forward_backward(data, fps_start=starts)
    o2 = graphed.forward_backward(None, fps_start=starts)
    assert rel_err(o2["total"], o1["total"]) <= 1e-6
    # fp32 atomics (weight gradients, scatter-adds) land in a different order every run, and a ReLU / arg-max flip moves
    # single entries: compare in relative L2 (two eager runs differ by the same amount)
    diff = (graphed.flat_grad - eager.flat_grad).double().norm() / eager.flat_grad.double().norm()
    assert float(diff) <= 1e-3, float(diff)
    for net in nets:
        net.train()
    graphed = GraphedTrainer(nets[1], data, lr=1e-3)               # train mode changed: a new capture is required
    losses = [float(graphed.step(data, fps_start=starts)["total"]) for _ in range(8)]
    assert losses[-1] < losses[0]


def test_pipelined_trainer_matches_eager(monkeypatch):
    """Two-stage pipelined training step (coordinate stage of batch i+1 beside the step of batch i, two buffer slots):
    batch by batch the loss and the gradients of the eager Trainer on the same batches in the same order (same CPU
    generator stream for the first FPS centroids), BatchNorm buffers untouched by the capture; then it trains."""
    from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
    from point2cyl_b200.train import PipelinedTrainer, Trainer
    B, N, K = 4, 2048, 4
    batches = [{k: v.to(DEV) for k, v in synthetic.s_cyl(B, N, K, seed=20 + i).items()} for i in range(5)]
    mask = ((torch.rand(B, 128, N, generator=torch.Generator().manual_seed(1)) > 0.5).float() * 2.0).to(DEV)
    monkeypatch.setattr(pipeline, "dropout_mask_fn", lambda x, p=0.5, **kw: mask)
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        nets.append(backbone(output_sizes=[3, 2 * K]).to(DEV).eval())   # running-statistics BatchNorm: well conditioned
    eager = Trainer(nets[0], lr=0.0)                       # lr 0: the weights stay equal, every batch is comparable
    piped = PipelinedTrainer(nets[1], batches[0], lr=0.0)
    for (k, a), (_, b) in zip(nets[0].named_buffers(), nets[1].named_buffers()):
        assert torch.equal(a, b), k
    torch.manual_seed(77)
    ref = []
    for b in batches:
        o = eager.step(b)
        ref.append((float(o["total"]), eager.flat_grad.clone()))
    torch.manual_seed(77)
    piped.prime(batches[0])
    for i in range(len(batches)):
        o = piped.step(batches[i + 1] if i + 1 < len(batches) else None)
        assert abs(float(o["total"]) - ref[i][0]) <= 1e-5 * max(1.0, abs(ref[i][0])), (i, float(o["total"]), ref[i][0])
        diff = (piped.flat_grad - ref[i][1]).double().norm() / ref[i][1].double().norm()
        assert float(diff) <= 1e-3, (i, float(diff))
    for net in nets:
        net.train()
    piped = PipelinedTrainer(nets[1], batches[0], lr=1e-3)
    piped.prime(batches[0])
    losses = [float(piped.step(None)["total"]) for _ in range(8)]  # step(None): the resident batches of the two slots
    assert losses[-1] < losses[0]


def _l2(got, ref):
    """relative L2 error: isolated ReLU / arg-max flips move single entries, not the bulk"""
    return float((got.detach().cpu().double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))


def _module_sd(sd, prefix):
    return {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


def test_module_level_autograd_sa_and_fp(monkeypatch):
    """Stand-alone PointNetSetAbstraction / PointNetFeaturePropagation drop-in modules are trainable: gradients w.r.t.
    their input features and parameters against torch autograd through the oracle (BatchNorm on running statistics,
    the well-conditioned setting)."""
    from point2cyl_b200.dropin.models import pointnet_util as pu
    B, N, K = 2, 1024, 4
    sd = orc.init_state_dict((3, 2 * K), seed=4)
    g = torch.Generator().manual_seed(8)
    xyz = synthetic.s_cyl(B, N, K, seed=4)["pcs"]
    # ---- sa2-shaped level: 512 points with 128 features -> 128 centres ----
    sa_spec = orc.SA_SPECS[1]
    xyz1 = xyz[:, :512].contiguous()
    feats = torch.randn(B, 128, 512, generator=g)
    start = torch.tensor([3, 7])
    monkeypatch.setattr(pipeline, "draw_fps_start", lambda B_, N_, dev: start.to(dev))
    sd_ref = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in sd.items()}
    f_ref = feats.clone().requires_grad_(True)
    new_xyz_ref, out_ref = orc.set_abstraction(sd_ref, sa_spec, xyz1.permute(0, 2, 1), f_ref, training=False, start=start)
    dO = torch.randn(out_ref.shape, generator=g)
    (out_ref * dO).sum().backward()
    sa = pu.PointNetSetAbstraction(sa_spec["npoint"], sa_spec["radius"], sa_spec["nsample"], 128 + 3, sa_spec["mlp"], False)
    sa.load_state_dict(_module_sd(sd, "sa2"), strict=True)
    sa = sa.to(DEV).eval()
    f_dev = feats.to(DEV).requires_grad_(True)
    new_xyz, out = sa(xyz1.permute(0, 2, 1).to(DEV), f_dev)
    assert rel_err(out, out_ref) <= TOL and rel_err(new_xyz, new_xyz_ref) == 0.0
    (out * dO.to(DEV)).sum().backward()
    assert _l2(f_dev.grad, f_ref.grad) <= 5e-3
    for k, p in sa.named_parameters():
        ref = sd_ref["sa2." + k].grad
        d = (p.grad.cpu().double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)
        assert float(d) <= 5e-3, (k, float(d))
    # ---- fp2-shaped level: 512 query points (128 skip features) <- 128 source points with 256 features ----
    fp_spec = orc.FP_SPECS[1]
    xyz2 = xyz[:, 512:640].contiguous()
    p1 = torch.randn(B, 128, 512, generator=g)
    p2 = torch.randn(B, 256, 128, generator=g)
    p1r, p2r = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
    for v in sd_ref.values():
        if v.is_floating_point() and v.grad is not None:
            v.grad = None
    o_ref = orc.feature_propagation(sd_ref, fp_spec, xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1), p1r, p2r, training=False)
    dO2 = torch.randn(o_ref.shape, generator=g)
    (o_ref * dO2).sum().backward()
    fp = pu.PointNetFeaturePropagation(384, fp_spec["mlp"])
    fp.load_state_dict(_module_sd(sd, "fp2"), strict=True)
    fp = fp.to(DEV).eval()
    p1d, p2d = p1.to(DEV).requires_grad_(True), p2.to(DEV).requires_grad_(True)
    o = fp(xyz1.permute(0, 2, 1).to(DEV), xyz2.permute(0, 2, 1).to(DEV), p1d, p2d)
    assert rel_err(o, o_ref) <= TOL
    (o * dO2.to(DEV)).sum().backward()
    assert _l2(p1d.grad, p1r.grad) <= 5e-3 and _l2(p2d.grad, p2r.grad) <= 5e-3
    for k, p in fp.named_parameters():
        ref = sd_ref["fp2." + k].grad
        d = (p.grad.cpu().double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)
        assert float(d) <= 5e-3, (k, float(d))


def test_unfused_set_abstraction_backward(monkeypatch):
    """A set-abstraction level whose first width is not 64 / 128 runs the materialised grouping path; its backward
    (p2c_group_bwd scatter + plain stack) against oracle autograd, BatchNorm on running statistics."""
    from point2cyl_b200.dropin.models import pointnet_util as pu
    g = torch.Generator().manual_seed(21)
    B, N, D = 2, 600, 16
    spec = dict(name="sax", npoint=64, radius=0.35, nsample=16, mlp=(32, 48), group_all=False)
    xyz = torch.rand(B, N, 3, generator=g)
    feats = torch.randn(B, D, N, generator=g)
    sd, cin = {}, 3 + D
    for i, cout in enumerate(spec["mlp"]):
        sd[f"sax.mlp_convs.{i}.weight"] = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
        sd[f"sax.mlp_convs.{i}.bias"] = torch.randn(cout, generator=g) * 0.1
        sd[f"sax.mlp_bns.{i}.weight"] = 1 + 0.2 * torch.randn(cout, generator=g)
        sd[f"sax.mlp_bns.{i}.bias"] = 0.1 * torch.randn(cout, generator=g)
        sd[f"sax.mlp_bns.{i}.running_mean"] = 0.1 * torch.randn(cout, generator=g)
        sd[f"sax.mlp_bns.{i}.running_var"] = 0.5 + torch.rand(cout, generator=g)
        sd[f"sax.mlp_bns.{i}.num_batches_tracked"] = torch.tensor(0)
        cin = cout
    start = torch.tensor([1, 5])
    monkeypatch.setattr(pipeline, "draw_fps_start", lambda B_, N_, dev: start.to(dev))
    sd_ref = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in sd.items()}
    f_ref = feats.clone().requires_grad_(True)
    _, out_ref = orc.set_abstraction(sd_ref, spec, xyz.permute(0, 2, 1), f_ref, training=False, start=start)
    dO = torch.randn(out_ref.shape, generator=g)
    (out_ref * dO).sum().backward()
    sa = pu.PointNetSetAbstraction(spec["npoint"], spec["radius"], spec["nsample"], 3 + D, list(spec["mlp"]), False)
    sa.load_state_dict(_module_sd(sd, "sax"), strict=True)
    sa = sa.to(DEV).eval()
    f_dev = feats.to(DEV).requires_grad_(True)
    _, out = sa(xyz.permute(0, 2, 1).to(DEV), f_dev)
    assert rel_err(out, out_ref) <= TOL
    (out * dO.to(DEV)).sum().backward()
    assert _l2(f_dev.grad, f_ref.grad) <= 5e-3
    for k, p in sa.named_parameters():
        assert _l2(p.grad, sd_ref["sax." + k].grad) <= 5e-3, k
