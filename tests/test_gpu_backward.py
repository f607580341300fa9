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
