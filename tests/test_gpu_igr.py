"""The implicit sketch network on the kernels against the reference goldens tests/golden/igr_*.npz
(produced by the reference's own IGR/network.py, IGR/sampler.py and the loss lines train_Point2Cyl.py:608-672) and
against the CPU oracle (oracle/igr_oracle.py, itself pinned to those goldens).  Bar: 1e-4 (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import igr_oracle as orc
from point2cyl_b200 import igr, ops
from point2cyl_b200.dropin.IGR import network as dnet
from point2cyl_b200.dropin.IGR.sampler import NormalPerPoint

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4
CASES = ["igr_b3_k2_s64.npz", "igr_b2_k4_s128_l2.npz"]


def rel_err(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a.detach().cpu()).double()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b.detach().cpu()).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def nets(seed):
    net = dnet.ImplicitNet(d_in=258, dims=[512] * 8, skip_in=[4], geometric_init=True, radius_init=1, beta=100)
    net.load_state_dict(orc.implicit_init(seed=seed), strict=True)
    enc = dnet.PointNetEncoder(256, 2, with_normals=True)
    enc.load_state_dict(orc.encoder_init(seed=seed + 1), strict=True)
    enc_gt = dnet.PointNetEncoder(256, 2, with_normals=True)
    enc_gt.load_state_dict(orc.encoder_init(seed=seed + 2), strict=True)
    return net.to(DEV), enc.to(DEV).train(), enc_gt.to(DEV).train()


def test_linear_act_epilogues_against_fp64():
    """p2c_linear_act on its own: softplus / sigmoid epilogue and multiplier epilogue, ragged N and K (254), strided
    outputs."""
    g = torch.Generator().manual_seed(0)
    for M, N, K in [(1000, 512, 512), (300, 254, 512), (513, 512, 254), (2048, 512, 258)]:
        ld = ops.pad4(K)
        X = torch.zeros(M, ld)
        X[:, :K] = torch.randn(M, K, generator=g) * 0.05
        W = torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g) * 0.01
        Xd, Wd = X.to(DEV), W.to(DEV)
        ws = ops.split_tf32_multi([Wd])[0]
        Z = X[:, :K].double() @ W.double().t() + b.double()
        wide = torch.full((M, ops.pad4(N) + 8), 7.0, device=DEV)
        S = torch.empty(M, ops.pad4(N), device=DEV)
        ops.linear_act(Xd, ws, b.to(DEV), N, K, op=1, beta=100.0, oscale=0.5, out=wide[:, 4:4 + N] if N % 4 == 0 else wide[:, :N],
                       S=S[:, :N])
        got = wide[:, 4:4 + N] if N % 4 == 0 else wide[:, :N]
        assert rel_err(got, 0.5 * torch.nn.functional.softplus(Z, beta=100.0)) <= 1e-5
        # (sigmoid(100 z) amplifies the 3xTF32 error of z by up to 25)
        assert rel_err(S[:, :N], torch.where(100 * Z > 20, torch.ones_like(Z), torch.sigmoid(100 * Z))) <= 5e-5
        untouched = wide[:, 4 + N:] if N % 4 == 0 else wide[:, ops.pad4(N):]
        assert bool((untouched == 7.0).all())                      # channels >= N are clipped, neighbours untouched
        mul = torch.randn(M, N, generator=g)
        obuf = torch.zeros(M, ops.pad4(N), device=DEV)
        mbuf = torch.zeros(M, ops.pad4(N), device=DEV)
        mbuf[:, :N] = mul.to(DEV)
        out = ops.linear_act(Xd, ws, None, N, K, op=2, oscale=2.0 ** -0.5, out=obuf[:, :N], mul=mbuf[:, :N])
        assert rel_err(out, (X[:, :K].double() @ W.double().t()) * mul.double() * 2.0 ** -0.5) <= 1e-5
        assert bool((obuf[:, N:] == 0).all())
        # the transposed split used by the reverse sweep: (a W)[:, :n_h]
        n_h = K - 4
        wt = ops.split_tf32_multi([Wd[:, :n_h]], transposed=[True])[0]
        A = torch.randn(M, N, generator=g) * 0.1
        Ap = torch.zeros(M, ops.pad4(N))
        Ap[:, :N] = A
        rbuf = torch.zeros(M, ops.pad4(n_h), device=DEV)
        r = ops.linear_act(Ap.to(DEV)[:, :N], wt, None, n_h, N, op=0, out=rbuf[:, :n_h])
        assert rel_err(r, A.double() @ W[:, :n_h].double()) <= 1e-5


@pytest.mark.parametrize("name", CASES)
def test_encoder_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    B, K, S, seed, is_l2 = (int(v) for v in g["meta"])
    net, enc, enc_gt = nets(seed)
    latent = enc(torch.from_numpy(g["global_pc"]).to(DEV))
    assert rel_err(latent, g["latent"]) <= TOL
    sk = torch.from_numpy(g["gt_sketches"]).reshape(B * K, S, 4).to(DEV)
    assert rel_err(enc_gt(sk), g["latent_gt"]) <= TOL


@pytest.mark.parametrize("name", CASES)
def test_implicit_network_golden(golden_dir, name):
    """f and d f / d x through the module-level drop-ins exactly as train_Point2Cyl.py:612-622 calls them."""
    g = np.load(os.path.join(golden_dir, name))
    B, K, S, seed, is_l2 = (int(v) for v in g["meta"])
    net, enc, enc_gt = nets(seed)
    latent = torch.from_numpy(g["latent"]).to(DEV)
    sk = torch.from_numpy(g["gt_sketches"]).reshape(B * K, S, 4).to(DEV)
    off = torch.from_numpy(g["nonmnfld_pnts"]).reshape(B * K, S + S // 8, 2).to(DEV)
    sk_pnts = dnet.add_latent(sk[:, :, :2], latent)
    nonmnfld_pnts = dnet.add_latent(off, latent)
    ref_x = orc.add_latent(sk[:, :, :2].cpu(), latent.cpu())
    assert torch.equal(sk_pnts.cpu(), ref_x)
    sk_pnts.requires_grad_()
    nonmnfld_pnts.requires_grad_()
    sk_pred = net(sk_pnts)
    nonmnfld_pred = net(nonmnfld_pnts)
    mnfld_grad = dnet.gradient(sk_pnts, sk_pred)
    nonmnfld_grad = dnet.gradient(nonmnfld_pnts, nonmnfld_pred)
    assert rel_err(sk_pred, g["sk_pred"]) <= TOL
    assert rel_err(mnfld_grad, g["mnfld_grad"]) <= TOL
    assert rel_err(nonmnfld_grad, g["nonmnfld_grad"]) <= TOL
    # a forward without requires_grad keeps nothing and still equals the oracle
    with torch.no_grad():
        f2 = net(sk_pnts.detach())
    assert not hasattr(f2, "_p2c_igr") and rel_err(f2, g["sk_pred"]) <= TOL


@pytest.mark.parametrize("name", CASES)
def test_sketch_loss_block_golden(golden_dir, name):
    """train_Point2Cyl.py:608-672 in one call: encoder latents, sampler stream, the four loss terms."""
    g = np.load(os.path.join(golden_dir, name))
    B, K, S, seed, is_l2 = (int(v) for v in g["meta"])
    net, enc, enc_gt = nets(seed)
    sk = torch.from_numpy(g["gt_sketches"]).reshape(B * K, S, 4).to(DEV)
    latent = enc(torch.from_numpy(g["global_pc"]).to(DEV))
    latent_gt = enc_gt(sk)
    off = torch.from_numpy(g["nonmnfld_pnts"]).reshape(B * K, S + S // 8, 2).to(DEV)
    mask_gt = torch.from_numpy(g["mask_gt"]).to(DEV)
    out = igr.sketch_loss_block(net, latent, latent_gt, sk[:, :, :2], sk[:, :, 2:], off, mask_gt, bool(is_l2))
    for k in ("im_loss", "mnfld_loss", "grad_loss", "normals_loss", "latent_loss"):
        assert rel_err(out[k], g[k]) <= TOL, k
    assert rel_err(out["sk_pred"], g["sk_pred"]) <= TOL
    assert rel_err(out["mnfld_grad"], g["mnfld_grad"]) <= TOL and rel_err(out["nonmnfld_grad"], g["nonmnfld_grad"]) <= TOL


def test_sampler_shapes_and_stream():
    s = NormalPerPoint(1.8, 0.01)
    pc = torch.rand(5, 64, 2, device=DEV)
    torch.manual_seed(3)
    a = s.get_points(pc)
    torch.manual_seed(3)
    b = s.get_points(pc)
    assert a.shape == (5, 72, 2) and torch.equal(a, b)
    assert float((a[:, :64] - pc).abs().max()) < 0.1 and float(a[:, 64:].abs().max()) <= 1.8


def test_implicit_network_larger_batch_vs_oracle():
    """A with-sketch-sized slice (8 instances x 1024 + 1152 points = 17,408 rows): values and input gradients against
    the CPU oracle's closed form, including rows on the linear branch of the softplus."""
    net, _, _ = nets(5)
    g = torch.Generator().manual_seed(6)
    I, S = 8, 1024
    latent = torch.nn.functional.normalize(torch.randn(I, 256, generator=g), dim=1)
    on = torch.rand(I, S, 2, generator=g) * 2 - 1
    off = torch.rand(I, S + S // 8, 2, generator=g) * 3.6 - 1.8
    f, ctx = igr.implicit_forward(net, latent=latent.to(DEV), pts=[on.to(DEV), off.to(DEV)])
    gx = igr.implicit_input_gradient(ctx)
    sd = orc.implicit_init(seed=5)
    x = torch.cat([orc.add_latent(on, latent), orc.add_latent(off, latent)])
    f_ref, g_ref = orc.implicit_forward_with_input_grad(sd, x)
    assert rel_err(f, f_ref) <= TOL and rel_err(gx, g_ref) <= TOL


# ---- second-order backward: training through the input gradient (train_Point2Cyl.py:608-672 + backward()) ----------


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().cpu().double().reshape(-1)
    b = torch.as_tensor(b).detach().cpu().double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_linear_act_bwd_epilogues_against_fp64():
    """p2c_linear_act_bwd on its own: op 3 (two outputs) and op 4 (addend, in place), ragged N / K (254)."""
    g = torch.Generator().manual_seed(1)
    for M, N, K in [(1000, 512, 512), (300, 254, 512), (513, 512, 254), (2048, 512, 258)]:
        X = torch.zeros(M, ops.pad4(K))
        X[:, :K] = torch.randn(M, K, generator=g) * 0.05
        W = torch.randn(N, K, generator=g) / K ** 0.5
        mul, v = torch.rand(M, N, generator=g), torch.randn(M, N, generator=g)
        pad = lambda t: torch.nn.functional.pad(t, (0, ops.pad4(N) - N)).to(DEV)
        ws = ops.split_tf32_multi([W.to(DEV)])[0]
        acc = X[:, :K].double() @ W.double().t()
        Y = torch.full((M, ops.pad4(N)), 7.0, device=DEV)
        Z = torch.full((M, ops.pad4(N)), 7.0, device=DEV)
        ops.linear_act_bwd(X.to(DEV)[:, :K], ws, N, K, 3, pad(mul)[:, :N], pad(v)[:, :N], Y[:, :N], beta=100.0,
                           oscale=2.0 ** -0.5, Z=Z[:, :N])
        assert rel_err(Y[:, :N], acc * mul.double() * 2.0 ** -0.5) <= 1e-5
        assert rel_err(Z[:, :N], 100.0 * acc * v.double() * (1.0 - mul.double())) <= 1e-5
        assert bool((Y[:, N:] == 7.0).all()) and bool((Z[:, N:] == 7.0).all())
        buf = pad(v).clone()
        ops.linear_act_bwd(X.to(DEV)[:, :K], ws, N, K, 4, pad(mul)[:, :N], buf[:, :N], buf[:, :N], oscale=0.5)
        assert rel_err(buf[:, :N], acc * mul.double() * 0.5 + v.double()) <= 1e-5


@pytest.mark.parametrize("which", ["both", "f_only", "g_only"])
def test_implicit_backward_vs_closed_form_oracle(which):
    """implicit_backward (kernels) against oracle.implicit_backward_closed_form evaluated in float64 on the same
    inputs and upstream gradients: every weight / bias gradient and d L / d x.  The oracle's closed form is itself
    checked against torch's double backward in tests/test_oracle_igr.py."""
    net, _, _ = nets(7)
    g = torch.Generator().manual_seed(8)
    I, S = 8, 128
    latent = torch.nn.functional.normalize(torch.randn(I, 256, generator=g), dim=1)
    on = torch.rand(I, S, 2, generator=g) * 2 - 1
    off = torch.rand(I, S + S // 8, 2, generator=g) * 3.6 - 1.8
    f, ctx = igr.implicit_forward(net, latent=latent.to(DEV), pts=[on.to(DEV), off.to(DEV)])
    igr.implicit_input_gradient(ctx)
    R = ctx.R
    f_bar = torch.randn(R, 1, generator=g) / R if which != "g_only" else None
    g_bar = torch.randn(R, 2, generator=g) / R if which != "f_only" else None
    grads, dx = igr.implicit_backward(ctx, None if f_bar is None else f_bar.to(DEV), None if g_bar is None else g_bar.to(DEV))
    sd = {k: v.double() for k, v in orc.implicit_init(seed=7).items()}
    x = torch.cat([orc.add_latent(on, latent), orc.add_latent(off, latent)]).double()
    ref_g, ref_dx = orc.implicit_backward_closed_form(
        sd, x, torch.zeros(R, 1).double() if f_bar is None else f_bar.double(),
        torch.zeros(R, 2).double() if g_bar is None else g_bar.double())
    named = dict(net.named_parameters())
    for k, ref in ref_g.items():
        assert rel_l2(grads[named[k]], ref) <= TOL, (k, rel_l2(grads[named[k]], ref))
    assert rel_l2(dx, ref_dx) <= TOL
    dlat = igr.latent_grad(ctx, dx)
    ref_dlat = ref_dx[:I * S, :256].reshape(I, S, 256).sum(1) + ref_dx[I * S:, :256].reshape(I, S + S // 8, 256).sum(1)
    assert rel_l2(dlat, ref_dlat) <= TOL


def _golden_grad_check(prefix, module, g, tol):
    """Sampled entries (first 2048) and L2 norm of every parameter gradient against the reference's.  A conv bias in
    front of a train-mode BatchNorm has an exactly zero gradient (the batch mean removes it): the reference's value is
    float32 rounding noise there, so such entries are compared on the scale of the module's largest gradient."""
    names = [n for n, _ in module.named_parameters()]
    top = max(float(g[f"gradnorm_{prefix}.{n}"]) for n in names)
    for pname, p in module.named_parameters():
        want, nrm = g[f"grad_{prefix}.{pname}"], float(g[f"gradnorm_{prefix}.{pname}"])
        assert p.grad is not None, (prefix, pname)
        got = p.grad.detach().reshape(-1)
        if nrm <= 1e-5 * top:
            assert float(got.double().norm()) <= 1e-4 * top, (prefix, pname, float(got.double().norm()), top)
            continue
        e_n = abs(float(got.double().norm()) - nrm) / nrm
        e_s = rel_l2(got[:want.size], want) if float(np.linalg.norm(want)) > 1e-3 * nrm else 0.0
        assert e_n <= tol and e_s <= tol, (prefix, pname, e_n, e_s)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("literal", [False, True])
def test_sketch_training_step_golden(golden_dir, name, literal):
    """im_loss.backward() on the drop-in modules against the gradients the REFERENCE produced for the same step
    (tests/golden/igr_*.npz: every parameter of ImplicitNet, of the trained encoder and of the gt encoder, plus
    d im_loss / d latent codes).  literal: the reference's own loss lines (train_Point2Cyl.py:608-672, read from the staged
    baseline/_ref file) exec'd on the drop-ins - value node + gradient() node; otherwise igr.sketch_loss_block (one
    fused node)."""
    g = np.load(os.path.join(golden_dir, name))
    B, K, S, seed, is_l2 = (int(v) for v in g["meta"])
    net, enc, enc_gt = nets(seed)
    sk = torch.from_numpy(g["gt_sketches"]).reshape(B * K, S, 4).to(DEV)
    mask_gt = torch.from_numpy(g["mask_gt"]).to(DEV)
    latent_codes = enc(torch.from_numpy(g["global_pc"]).to(DEV))
    latent_codes.retain_grad()
    sk_pnts, sk_normals = sk[:, :, :2].contiguous(), sk[:, :, 2:].contiguous()
    latent_codes_gt = enc_gt(torch.cat((sk_pnts, sk_normals), dim=-1))
    off = torch.from_numpy(g["nonmnfld_pnts"]).reshape(B * K, S + S // 8, 2).to(DEV)
    if literal:
        from baseline import ref_arm
        root = ref_arm.reference_root()
        if root is None or not os.path.exists(os.path.join(root, "train_Point2Cyl.py")):
            pytest.skip("baseline/_ref not staged")
        import textwrap
        lines = open(os.path.join(root, "train_Point2Cyl.py")).readlines()[607:672]
        src = textwrap.dedent("".join(l.replace("\t", "    ") for l in lines))
        assert src.startswith("if WITH_IM_LOSS:") and src.rstrip().endswith("im_loss += latent_loss")
        from point2cyl_b200.dropin.losses import reduce_mean_masked_instance

        class FixedSampler:                      # the golden's off-surface sample (the reference drew it on the CPU)
            def get_points(self, pc):
                return off

        ns = dict(torch=torch, sampler=FixedSampler(), add_latent=dnet.add_latent, gradient=dnet.gradient,
                  implicit_net=net, reduce_mean_masked_instance=reduce_mean_masked_instance, mask_gt=mask_gt,
                  sk_pnts=sk_pnts, sk_normals=sk_normals, latent_codes=latent_codes, latent_codes_gt=latent_codes_gt,
                  batch_size=B, K=K, WITH_IM_LOSS=True, IS_L2=bool(is_l2), pcs=sk)
        exec(src, ns)
        im_loss = ns["im_loss"]
    else:
        im_loss = igr.sketch_loss_block(net, latent_codes, latent_codes_gt, sk_pnts, sk_normals, off, mask_gt,
                                        bool(is_l2))["im_loss"]
    assert rel_err(im_loss, g["im_loss"]) <= TOL
    im_loss.backward()
    assert rel_l2(latent_codes.grad, g["d_latent"]) <= TOL
    # measured (tests/tools/igr_grad_err.py): net <= 3.0e-5, encoders <= 2.3e-5, d latent <= 1.1e-5
    _golden_grad_check("net", net, g, TOL)
    _golden_grad_check("enc", enc, g, TOL)
    _golden_grad_check("encgt", enc_gt, g, TOL)
