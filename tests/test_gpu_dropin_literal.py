"""The literal drop-in proof.

The forward+loss statements of the reference's training loop - `X, W_raw = model(pcs)` ... `total_loss +=
total_center_loss`, train_Point2Cyl_without_sketch.py:244-353 - are read from the reference's OWN file (staged under
the git-ignored baseline/_ref/ by baseline/make_ref.py; nothing of it is committed here) and exec'd UNCHANGED with

    model                        = point2cyl_b200.dropin.models.pointnet_extrusion.backbone
    compute_all_losses, ...      = point2cyl_b200.dropin.losses
    estimate_extrusion_axis, ... = point2cyl_b200.dropin.data_utils

bound in place of the reference's modules; then `total_loss.backward(); optimizer.step()` as the script does
(:367-368).  Losses are compared with the reference's own run of the same lines (tests/golden/train_*.npz) and, when
the reference modules are importable (baseline/_ref), with the SAME exec'd lines bound to the reference on the CPU.
Skipped when the staged reference files are absent.
"""
import os

import numpy as np
import pytest
import torch

from oracle import p2c_oracle as orc
from point2cyl_b200 import pipeline, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


def rel_err(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _dropin_bindings():
    from point2cyl_b200.dropin import data_utils as du
    from point2cyl_b200.dropin import losses as L
    return dict(compute_all_losses=L.compute_all_losses, get_mask_gt=L.get_mask_gt,
                compute_normal_loss=L.compute_normal_loss, reduce_mean_masked_instance=L.reduce_mean_masked_instance,
                estimate_extrusion_axis=du.estimate_extrusion_axis, estimate_extrusion_centers=du.estimate_extrusion_centers)


LOSS_NAMES = ("total_loss", "total_normal_loss", "total_miou_loss", "total_bb_loss", "total_extrusion_loss",
              "total_center_loss")


@pytest.mark.parametrize("name,training", [("train_bneval_b2_n1024_k4.npz", False), ("train_b2_n1024_k4.npz", True)])
def test_reference_loop_body_runs_on_the_dropin(golden_dir, name, training):
    from baseline import ref_arm
    from point2cyl_b200.dropin.models.pointnet_extrusion import backbone
    root = ref_arm.reference_root()
    if root is None:
        pytest.skip("baseline/_ref not staged (python baseline/make_ref.py where /root/reference exists)")
    g = np.load(os.path.join(golden_dir, name))
    B, N, K, seed = (int(v) for v in g["meta"])
    data = synthetic.s_cyl(B, N, K, seed)
    sd = orc.init_state_dict((3, 2 * K), seed=seed)
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0

    # ---- the drop-in, on the GPU ----
    model = backbone(output_sizes=[3, 2 * K])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).train(training)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)          # train_Point2Cyl_without_sketch.py:189
    body = ref_arm.LoopBody(root, model, _dropin_bindings(), K, N)
    dev = {k: v.to(DEV) for k, v in data.items()}
    mdev = mask.to(DEV)
    real = pipeline.dropout_mask_fn
    pipeline.dropout_mask_fn = lambda ones, p=0.5: mdev
    try:
        torch.manual_seed(seed)                                        # the FPS starts come from the CPU generator (:75)
        ns = body(dev)
    finally:
        pipeline.dropout_mask_fn = real
    assert np.array_equal(ns["matching_indices"].cpu().numpy(), g["matching_indices"])
    ftol = TOL if not training else 5e-4      # train-mode BatchNorm on a batch of 2: tests/test_gpu_fullsize.py adjudicates
    # (X is the NORMALISED prediction here - the script overwrites it, :247: short raw normals amplify the error)
    assert rel_err(ns["X"], torch.nn.functional.normalize(torch.from_numpy(g["X_raw"]), dim=2)) <= (TOL if not training else 1e-2)
    assert rel_err(ns["W_raw"], g["W_raw"]) <= ftol
    assert rel_err(ns["total_loss"], g["loss"]) <= ftol
    before = [p.detach().clone() for p in model.parameters()]
    optimizer.zero_grad()                                              # :355
    ns["total_loss"].backward()                                        # :367
    optimizer.step()                                                   # :368
    grads = [p.grad for p in model.parameters()]
    assert all(gr is not None and bool(torch.isfinite(gr).all()) for gr in grads)
    assert sum(float((p.detach() - q).abs().max() > 0) for p, q in zip(model.parameters(), before)) >= len(before) - 8

    # ---- the same exec'd lines bound to the reference's own modules, on the CPU ----
    ref = ref_arm.load_modules(root)
    rmodel = ref.net.backbone(output_sizes=[3, 2 * K])
    rmodel.load_state_dict(sd, strict=True)
    rmodel.train(training)
    rbody = ref_arm.LoopBody(root, rmodel, ref_arm.reference_bindings(ref), K, N)
    real_dropout = ref.net.F.dropout
    ref.net.F.dropout = lambda x, p=0.5, **kw: x * mask
    try:
        torch.manual_seed(seed)
        rns = rbody(data)
    finally:
        ref.net.F.dropout = real_dropout
    assert rel_err(rns["total_loss"], g["loss"]) <= 1e-5               # the exec'd lines ARE what made the golden
    for k in LOSS_NAMES:
        assert rel_err(ns[k], rns[k]) <= ftol, k
    m = rns["mask_gt"]
    dots = (ns["E_AX"].detach().cpu() * rns["E_AX"].detach()).sum(-1).abs()
    assert float((1 - dots[m]).max()) <= ftol
    assert rel_err(ns["predicted_centroids"].detach().cpu()[m], rns["predicted_centroids"].detach()[m]) <= ftol
