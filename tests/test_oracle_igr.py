"""Pins oracle/igr_oracle.py (the implicit sketch network of the with-sketch trainer) to goldens
produced by the reference's own IGR modules and loss lines (tests/golden/make_golden.py:igr_case).  CPU only: there is
no CUDA path for this block yet - these tests fix the arithmetic it will have to match."""
import os

import numpy as np
import pytest
import torch

from oracle import igr_oracle as igr

CASES = ["igr_b3_k2_s64.npz", "igr_b2_k4_s128_l2.npz"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def rel_err(a, b):
    a = torch.as_tensor(np.asarray(a)).double() if not torch.is_tensor(a) else a.detach().double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def setup(g):
    B, K, S, seed, is_l2 = (int(v) for v in g["meta"])
    net = igr.implicit_init(seed=seed)
    enc = igr.encoder_init(seed=seed + 1)
    enc_gt = igr.encoder_init(seed=seed + 2)
    gt_sketches = torch.from_numpy(g["gt_sketches"])
    global_pc = torch.from_numpy(g["global_pc"])
    mask_gt = torch.from_numpy(g["mask_gt"])
    return B, K, S, seed, bool(is_l2), net, enc, enc_gt, gt_sketches, global_pc, mask_gt


def test_layer_shapes_follow_the_skip_rule():
    shapes = igr.implicit_layer_shapes()
    assert shapes[0] == (258, 512) and shapes[3] == (512, 254) and shapes[4] == (512, 512) and shapes[-1] == (512, 1)
    assert sum(fi * fo + fo for fi, fo in shapes) == 1_839_359        # parameters of the 8 x 512 network


@pytest.mark.parametrize("name", CASES)
def test_encoder_and_sampler(golden_dir, name):
    g = load(golden_dir, name)
    B, K, S, seed, is_l2, net, enc, enc_gt, gt_sketches, global_pc, mask_gt = setup(g)
    latent = igr.encoder_forward(enc, global_pc, training=True)
    assert rel_err(latent, g["latent"]) <= 1e-5
    assert torch.allclose(latent.norm(dim=1), torch.ones(B * K), atol=1e-5)      # F.normalize
    sk = gt_sketches.reshape(B * K, S, 4)
    assert rel_err(igr.encoder_forward(enc_gt, sk, training=True), g["latent_gt"]) <= 1e-5
    torch.manual_seed(seed + 3)
    off = igr.sample_off_surface(sk[:, :, :2])
    assert off.shape == (B * K, S + S // 8, 2)
    assert np.array_equal(off.reshape(-1, 2).numpy(), g["nonmnfld_pnts"])       # same generator stream, bit-exact


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("analytic", [False, True])
def test_sketch_loss_block(golden_dir, name, analytic):
    """Prediction, input gradient and the four loss terms; `analytic` = the closed-form input gradient a kernel will
    compute (forward keeping pre-activations, sigmoid(beta z) chain) instead of autograd."""
    g = load(golden_dir, name)
    B, K, S, seed, is_l2, net, enc, enc_gt, gt_sketches, global_pc, mask_gt = setup(g)
    sk = gt_sketches.reshape(B * K, S, 4)
    latent = torch.from_numpy(g["latent"])
    latent_gt = torch.from_numpy(g["latent_gt"])
    off = torch.from_numpy(g["nonmnfld_pnts"]).reshape(B * K, S + S // 8, 2)
    out = igr.sketch_loss_block(net, latent, latent_gt, sk[:, :, :2], sk[:, :, 2:], off, mask_gt, is_l2, analytic)
    assert rel_err(out["sk_pred"], g["sk_pred"]) <= 1e-5
    assert rel_err(out["mnfld_grad"], g["mnfld_grad"]) <= 1e-5
    assert rel_err(out["nonmnfld_grad"], g["nonmnfld_grad"]) <= 1e-5
    for k in ("im_loss", "mnfld_loss", "grad_loss", "normals_loss", "latent_loss"):
        assert rel_err(out[k], g[k]) <= 1e-5, k


@pytest.mark.parametrize("name", CASES)
def test_sketch_training_gradients(golden_dir, name):
    """One backward of im_loss through the double-differentiated network and both encoders reproduces the reference's
    parameter gradients (stored as the first 2048 entries + L2 norm) and d im_loss / d latent."""
    g = load(golden_dir, name)
    B, K, S, seed, is_l2, net, enc, enc_gt, gt_sketches, global_pc, mask_gt = setup(g)
    for sd in (net, enc, enc_gt):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    sk = gt_sketches.reshape(B * K, S, 4)
    latent = igr.encoder_forward(enc, global_pc, training=True)
    latent.retain_grad()
    latent_gt = igr.encoder_forward(enc_gt, sk, training=True)
    torch.manual_seed(seed + 3)
    off = igr.sample_off_surface(sk[:, :, :2])
    out = igr.sketch_loss_block(net, latent, latent_gt, sk[:, :, :2], sk[:, :, 2:], off, mask_gt, is_l2)
    out["im_loss"].backward()
    assert rel_err(out["im_loss"], g["im_loss"]) <= 1e-5
    assert rel_err(latent.grad, g["d_latent"]) <= 1e-4
    checked = 0
    for prefix, sd in (("net", net), ("enc", enc), ("encgt", enc_gt)):
        for k, v in sd.items():
            key = f"grad_{prefix}.{k}"
            if key not in g.files:
                continue
            ref = torch.from_numpy(g[key]).double()
            got = v.grad.reshape(-1).double()
            nref = float(g[f"gradnorm_{prefix}.{k}"])
            assert abs(float(got.norm()) - nref) <= 2e-4 * max(nref, 1e-6) + 1e-9, key
            scale = max(float(ref.abs().max()), 1e-3 * nref, 1e-12)
            assert float((got[:ref.numel()] - ref).abs().max()) <= 2e-4 * scale + 1e-9, key
            checked += 1
    assert checked == 18 + 22 + 22          # 9 Linear (w,b) + 2 x (5 conv (w,b) + 5 BN (w,b) + fc (w,b))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_closed_form_second_order_backward_matches_autograd(dtype, tol):
    """The four-sweep backward of (f, df/dx) written out without autograd - the blueprint for the kernels - equals
    torch's double backward, including rows on the linear branch of softplus (beta z > 20) and the skip layer."""
    sd = {k: v.to(dtype) for k, v in igr.implicit_init(seed=3).items()}
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(200, 258, generator=g) * 0.3).to(dtype)
    x[:40] *= 4.0
    f_bar = torch.randn(200, 1, generator=g).to(dtype)
    g_bar = torch.randn(200, 2, generator=g).to(dtype)
    P = {k: v.clone().requires_grad_() for k, v in sd.items()}
    xa = x.clone().requires_grad_()
    f = igr.implicit_forward(P, xa)
    gx = igr.gradient(xa, f)
    ((f * f_bar).sum() + (gx * g_bar).sum()).backward()
    grads, dx = igr.implicit_backward_closed_form(sd, x, f_bar, g_bar)
    assert set(grads) == set(sd)
    for k in sd:
        ref = P[k].grad
        assert float((grads[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-30)) <= tol, k
    assert float((dx - xa.grad).abs().max() / xa.grad.abs().max()) <= tol
