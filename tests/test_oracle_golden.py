"""Pins the oracle restatement (oracle/p2c_oracle.py) to the golden vectors produced by the upstream
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import p2c_oracle as orc
from point2cyl_b200 import synthetic

FLOAT_TOL = 1e-6


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


POINTOPS = [("pointops_cyl_n1024.npz", "cyl"), ("pointops_uniform_n2048.npz", "uniform")]


def pointops_inputs(g, kind):
    B, N, npoint, nsample, seed = (int(v) for v in g["meta"])
    xyz = synthetic.s_cyl(B, N, 4, seed)["pcs"] if kind == "cyl" else synthetic.s_uniform(B, N, seed)
    return xyz, npoint, float(g["radius"]), nsample, seed


@pytest.mark.parametrize("name,kind", POINTOPS)
def test_fps_and_ball_query_bit_exact(golden_dir, name, kind):
    g = load(golden_dir, name)
    xyz, npoint, radius, nsample, seed = pointops_inputs(g, kind)
    start = torch.from_numpy(g["start"])
    fps = orc.farthest_point_sample(xyz, npoint, start)
    assert np.array_equal(fps.numpy(), g["fps_idx"].astype(np.int64))
    new_xyz = orc.gather_points(xyz, fps)
    grp = orc.query_ball_point(radius, nsample, xyz, new_xyz)
    assert np.array_equal(grp.numpy(), g["group_idx"].astype(np.int64))
    d = orc.square_distance(new_xyz, xyz)
    assert np.array_equal(d[:, 0, :].numpy(), g["sqdist_row0"])


@pytest.mark.parametrize("name,kind", POINTOPS)
def test_fma_recipe_matches_matmul(golden_dir, name, kind):
    """The dot-product recipe the CUDA kernels implement == what torch.matmul produced in the goldens."""
    g = load(golden_dir, name)
    xyz, npoint, radius, nsample, seed = pointops_inputs(g, kind)
    fps = torch.from_numpy(g["fps_idx"].astype(np.int64))
    new_xyz = orc.gather_points(xyz, fps)
    d = orc.square_distance_fma_emulated(new_xyz[:, :1], xyz)
    assert np.array_equal(d[:, 0, :].numpy(), g["sqdist_row0"])


@pytest.mark.parametrize("name,kind", POINTOPS)
def test_three_nn_interp(golden_dir, name, kind):
    g = load(golden_dir, name)
    xyz, npoint, radius, nsample, seed = pointops_inputs(g, kind)
    B = xyz.shape[0]
    fps = torch.from_numpy(g["fps_idx"].astype(np.int64))
    new_xyz = orc.gather_points(xyz, fps)
    feats2 = torch.randn(B, npoint, 16, generator=torch.Generator().manual_seed(seed + 1))
    interp, idx, w = orc.three_nn_interpolate(xyz, new_xyz, feats2)
    assert np.array_equal(idx.numpy(), g["nn_idx"].astype(np.int64))
    assert np.array_equal(w.numpy(), g["nn_w"])
    assert rel_err(interp, g["interp"]) <= FLOAT_TOL


@pytest.mark.parametrize("name", ["backbone_b2_n1024_k4.npz", "backbone_b1_n1024_k4.npz"])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_backbone(golden_dir, name, mode):
    g = load(golden_dir, name)
    B, N, K, seed = (int(v) for v in g["meta"])
    data = synthetic.s_cyl(B, N, K, seed)
    sd = orc.init_state_dict(output_sizes=(3, 2 * K), seed=seed)
    starts = (torch.from_numpy(g[f"{mode}_s1"]), torch.from_numpy(g[f"{mode}_s2"]))
    new_stats = {}
    with torch.no_grad():
        X, W = orc.backbone_forward(sd, data["pcs"], training=(mode == "train"), fps_start=starts,
                                    new_stats=new_stats)
    assert rel_err(X, g[f"{mode}_X"]) <= FLOAT_TOL
    assert rel_err(W, g[f"{mode}_W"]) <= FLOAT_TOL
    if mode == "train":
        for k in g.files:
            if k.startswith("stat_") and "num_batches" not in k:
                assert rel_err(new_stats[k[5:]], g[k]) <= FLOAT_TOL, k


LOSS = ["loss_b2_n1024_k4.npz", "loss_b3_n2048_k8_normeig.npz"]


def loss_inputs(g):
    B, N, K, seed, norm_eig = (int(v) for v in g["meta"])
    data = synthetic.s_cyl(B, N, K, seed)
    return data, torch.from_numpy(g["X_raw"]), torch.from_numpy(g["W_raw"]), bool(norm_eig)


@pytest.mark.parametrize("name", LOSS)
@pytest.mark.parametrize("dense", [True, False])
def test_loss_block(golden_dir, name, dense):
    g = load(golden_dir, name)
    data, X_raw, W_raw, norm_eig = loss_inputs(g)
    if dense and X_raw.shape[1] > 1024:
        pytest.skip("dense (B,N,N) route only on the small case")
    out = orc.loss_block(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"],
                         data["axes"], data["centers"], norm_eig=norm_eig, dense_axis=dense)
    assert np.array_equal(out["matching_indices"].numpy(), g["matching_indices"])
    assert np.array_equal(out["mask"].numpy(), g["mask"])
    for k in ("normal", "miou", "bb", "center"):
        assert rel_err(out[k], g[k]) <= 2e-6, k
    # axes: sign-free comparison; the fp32 dense route and the f64 scatter route agree to ~1e-5
    dots = (out["E_AX"] * torch.from_numpy(g["E_AX"])).sum(-1).abs()
    m = torch.from_numpy(g["mask"])
    assert float((1 - dots[m]).max()) <= 1e-5
    assert rel_err(out["axis"], g["axis"]) <= 1e-4
    assert rel_err(out["centers"], g["centers"]) <= 2e-6


@pytest.mark.parametrize("name", LOSS)
def test_eval_helpers(golden_dir, name):
    g = load(golden_dir, name)
    data, X_raw, W_raw, _ = loss_inputs(g)
    X, Wb, Wc, W = orc.postnet_split(X_raw, W_raw)
    hard = orc.hard_W_encoding(W, to_null_mask=True)
    assert np.array_equal(hard.argmax(-1).numpy().astype(np.int8), g["hard_argmax"])
    assert np.array_equal(hard.sum(-1).numpy().astype(np.int8), g["hard_rowsum"])
    match = torch.from_numpy(g["matching_indices"])
    iou = orc.compute_segmentation_iou(W, data["inst"], match, torch.from_numpy(g["mask"]).float())
    assert rel_err(iou, g["seg_iou"]) <= 2e-6
    nd = orc.compute_normal_difference(X, data["normals"])
    assert rel_err(nd, g["normal_diff"]) <= 2e-6


def projection_inputs(g):
    B, N, K, S, seed = (int(v) for v in g["meta"])
    data = synthetic.s_cyl(B, N, K, seed)
    return (data["pcs"], data["normals"], torch.from_numpy(g["inst"]).long(), torch.from_numpy(g["bb"]).long(),
            torch.from_numpy(g["axes"]), torch.from_numpy(g["centers"]), S, seed)


def test_projection_closed_forms(golden_dir):
    """a19: sketch_implicit_projection{,2,3} and get_extrusion_extents of the reference (run with the restated
    torchgeometry rotation) vs the oracle, same CPU random stream."""
    g = load(golden_dir, "projection_b3_n512_k4.npz")
    P, X, inst, bb, axes, centers, S, seed = projection_inputs(g)
    torch.manual_seed(seed)
    Pp, Xp, sc, found = orc.sketch_implicit_projection(P, X, inst, bb, axes, centers, S)
    assert np.array_equal(found.numpy(), g["found"])
    assert found.numpy().min() == 0 and found.numpy().max() == 1      # both branches present
    assert rel_err(Pp, g["P_proj"]) <= FLOAT_TOL and rel_err(Xp, g["X_proj"]) <= FLOAT_TOL
    assert rel_err(sc, g["scales"]) <= FLOAT_TOL
    Pp3, Xp3, sc3, found3 = orc.sketch_implicit_projection(P, X, inst, bb, axes, centers, P.shape[1], all_points=True)
    assert np.array_equal(found3.numpy(), g["found3"])
    assert rel_err(Pp3, g["P_proj3"]) <= FLOAT_TOL and rel_err(Xp3, g["X_proj3"]) <= FLOAT_TOL
    assert rel_err(sc3, g["scales3"]) <= FLOAT_TOL
    torch.manual_seed(seed + 1)
    ext, found_e = orc.get_extrusion_extents(P, inst, bb, axes, centers, S)
    assert np.array_equal(found_e.numpy(), g["found_ext"])
    assert rel_err(ext, g["extents"]) <= FLOAT_TOL


def test_rotation_restatement_is_a_rotation_only_for_unit_axis():
    """Documents the reference quirk (data_utils.py:1096-1103): the rotation vector is (a x z)*angle with a x z NOT
    normalised, so the matrix rotates by angle*sin(angle), and only takes a to +z when the two coincide."""
    a = torch.nn.functional.normalize(torch.tensor([[0.3, -0.5, 0.6]]), dim=-1)
    z = torch.tensor([[0.0, 0.0, 1.0]])
    ang = torch.acos((a * z).sum(-1))
    R = orc.angle_axis_to_rotation_matrix(torch.linalg.cross(a, z) * ang[:, None])[0, :3, :3]
    assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5)            # still orthonormal
    got = torch.acos(((R @ R @ R).trace() - 1) / 2)                    # 3 * rotation angle
    assert abs(float(got) - 3 * float(ang * torch.sin(ang))) < 1e-4


def test_rotation_restatement_against_scipy_rodrigues():
    """torchgeometry is not installable offline, so the restated angle_axis_to_rotation_matrix is anchored on an
    independent implementation of the same published formula (Rodrigues): scipy's Rotation.from_rotvec.  The only
    difference is torchgeometry's `theta + 1e-6` in the axis normalisation (<= 2e-6 relative)."""
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(11)
    aa = torch.randn(256, 3, generator=g, dtype=torch.float64) * torch.rand(256, 1, generator=g, dtype=torch.float64) * 3.0
    R = orc.angle_axis_to_rotation_matrix(aa)[:, :3, :3]
    ref = torch.from_numpy(Rotation.from_rotvec(aa.numpy()).as_matrix())
    assert float((R - ref).abs().max()) <= 1e-5
    tiny = torch.randn(64, 3, generator=g, dtype=torch.float64) * 1e-4          # first-order branch (theta^2 <= 1e-6)
    Rt = orc.angle_axis_to_rotation_matrix(tiny)[:, :3, :3]
    reft = torch.from_numpy(Rotation.from_rotvec(tiny.numpy()).as_matrix())
    assert float((Rt - reft).abs().max()) <= 1e-7                               # second-order terms only
    assert torch.equal(orc.angle_axis_to_rotation_matrix(aa)[:, 3, :], torch.tensor([0, 0, 0, 1.0], dtype=torch.float64).expand(256, 4))


@pytest.mark.parametrize("name", LOSS)
def test_loss_gradients(golden_dir, name):
    """Autograd through the oracle's loss block reproduces the reference's own gradients w.r.t. the network
    outputs (pins the checker used for the backward kernels)."""
    g = load(golden_dir, name)
    data, X_raw, W_raw, norm_eig = loss_inputs(g)
    X_raw = X_raw.clone().requires_grad_(True)
    W_raw = W_raw.clone().requires_grad_(True)
    out = orc.loss_block(data["pcs"], X_raw, W_raw, data["normals"], data["inst"], data["bb"],
                         data["axes"], data["centers"], norm_eig=norm_eig)
    out["total"].backward()
    assert rel_err(X_raw.grad, g["dX_raw"]) <= 2e-5
    assert rel_err(W_raw.grad, g["dW_raw"]) <= 2e-5


def grad_errors(got, g, key):
    """(relative L2 error, median |err| / max|ref|, norm error) of a parameter gradient against its stored sample
    (first 2048 entries + L2 norm)."""
    ref = torch.from_numpy(g["grad_" + key]).double()
    flat = got.detach().reshape(-1).double().cpu()
    d = flat[:ref.numel()] - ref
    scale = max(float(ref.abs().max()), 1e-30)
    return (float(d.norm() / ref.norm().clamp_min(1e-30)), float(d.abs().median()) / scale,
            abs(float(flat.norm()) - float(g["gnorm_" + key])) / max(float(g["gnorm_" + key]), 1e-30))


def live_keys(g, training):
    """Parameters whose gradient is not mathematically zero: with train-mode BatchNorm a conv bias in front of it,
    or a shift that a later BatchNorm removes (sa3's last beta), only carries rounding noise."""
    gmax = max(float(g[k]) for k in g.files if k.startswith("gnorm_"))
    keys = [k[5:] for k in g.files if k.startswith("grad_") and float(g["gnorm_" + k[5:]]) > 1e-5 * gmax]
    if training:
        keys = [k for k in keys if not (k.endswith("bias") and ("mlp_convs" in k or k.startswith("fc1")))]
    return keys


@pytest.mark.parametrize("name,training", [("train_b2_n1024_k4.npz", True), ("train_bneval_b2_n1024_k4.npz", False)])
def test_training_step_gradients(golden_dir, name, training):
    """Autograd through the oracle's backbone + loss block reproduces the reference's parameter gradients for one
    training step (pins the checker of the backward kernels).  Same torch ops in the same order, so this holds to
    1e-3 even though the step is ill-conditioned (see test_gpu_backward.py)."""
    g = load(golden_dir, name)
    B, N, K, seed = (int(v) for v in g["meta"])
    data = synthetic.s_cyl(B, N, K, seed)
    sd = orc.init_state_dict((3, 2 * K), seed=seed)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd.items()}
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0
    starts = (torch.from_numpy(g["s1"]), torch.from_numpy(g["s2"]))
    out = orc.forward_loss(sd, data, training=training, fps_start=starts, dropout_mask=mask)
    assert rel_err(out["total"].detach(), g["loss"]) <= 1e-5
    out["total"].backward()
    assert len([k for k in g.files if k.startswith("grad_")]) == 72
    keys = live_keys(g, training)
    assert len(keys) >= 40
    for k in keys:
        l2, med, nerr = grad_errors(sd[k].grad, g, k)
        assert l2 <= 1e-3 and nerr <= 1e-3, (k, l2, med, nerr)
