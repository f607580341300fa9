"""Generate the golden vectors under tests/golden/ by running the UPSTREAM reference.

Run in the authoring container (the reference tree is at /root/reference there):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures of its own, so these files are the
pin for the oracle restatement (oracle/p2c_oracle.py) and, through it, for the CUDA path.  Inputs
are regenerated deterministically from seeds by point2cyl_b200.synthetic / oracle.init_state_dict,
so only the reference OUTPUTS are stored.  The training script's inline base/barrel loss
(train_Point2Cyl_without_sketch.py:283-313) is not an importable function; its source lines are
read from the reference file and exec'd here so that the golden is the reference's own code.
"""
import os
import sys
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import p2c_oracle as orc  # noqa: E402
from oracle import igr_oracle as igr  # noqa: E402
from point2cyl_b200 import synthetic  # noqa: E402


def np32(t):
    return t.detach().cpu().numpy().copy()  # copy: state_dict tensors are live views


def pointops_case(ref, name, xyz, npoint, radius, nsample, seed):
    """FPS / ball query / gather / 3-NN interpolation known answers."""
    B, N, _ = xyz.shape
    torch.manual_seed(seed)
    start = torch.randint(0, N, (B,), dtype=torch.long)
    torch.manual_seed(seed)
    fps_idx = ref.util.farthest_point_sample(xyz, npoint)
    assert torch.equal(fps_idx[:, 0], start)
    new_xyz = ref.util.index_points(xyz, fps_idx)
    grp = ref.util.query_ball_point(radius, nsample, xyz, new_xyz)
    d = ref.util.square_distance(new_xyz, xyz)
    # three-NN interpolation exactly as PointNetFeaturePropagation.forward does it (:301-308)
    g = torch.Generator().manual_seed(seed + 1)
    feats2 = torch.randn(B, npoint, 16, generator=g)
    dists, idx = ref.util.square_distance(xyz, new_xyz).sort(dim=-1)
    dists, idx = dists[:, :, :3], idx[:, :, :3]
    recip = 1.0 / (dists + 1e-8)
    w = recip / torch.sum(recip, dim=2, keepdim=True)
    interp = torch.sum(ref.util.index_points(feats2, idx) * w.view(B, N, 3, 1), dim=2)
    np.savez_compressed(
        os.path.join(HERE, name),
        start=np32(start), fps_idx=np32(fps_idx).astype(np.int32), group_idx=np32(grp).astype(np.int32),
        sqdist_row0=np32(d[:, 0, :]), nn_idx=np32(idx).astype(np.int32), nn_w=np32(w), interp=np32(interp),
        meta=np.array([B, N, npoint, nsample, seed], dtype=np.int64), radius=np.float64(radius))
    print("wrote", name)


def backbone_case(ref, name, B, N, K, seed):
    data = synthetic.s_cyl(B, N, K, seed)
    sd = orc.init_state_dict(output_sizes=(3, 2 * K), seed=seed)
    net = ref.net.backbone(output_sizes=[3, 2 * K])
    net.load_state_dict(sd, strict=True)
    out = {}
    real_dropout = ref.net.F.dropout
    ref.net.F.dropout = lambda x, p=0.5, **kw: x  # identity on both sides
    try:
        for mode in ("train", "eval"):
            net.load_state_dict(sd, strict=True)
            net.train(mode == "train")
            torch.manual_seed(seed)
            s1 = torch.randint(0, N, (B,), dtype=torch.long)
            s2 = torch.randint(0, 512, (B,), dtype=torch.long)
            torch.manual_seed(seed)
            with torch.no_grad():
                X, W = net(data["pcs"])
            out[f"{mode}_X"] = np32(X)
            out[f"{mode}_W"] = np32(W)
            out[f"{mode}_s1"] = np32(s1)
            out[f"{mode}_s2"] = np32(s2)
            if mode == "train":
                new = net.state_dict()
                for k in ("sa1.mlp_bns.0.running_mean", "sa1.mlp_bns.0.running_var",
                          "sa3.mlp_bns.2.running_var", "fp1.mlp_bns.2.running_mean", "bn1.running_var",
                          "bn1.num_batches_tracked"):
                    out["stat_" + k] = np32(new[k])
    finally:
        ref.net.F.dropout = real_dropout
    out["meta"] = np.array([B, N, K, seed], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name)


def _inline_bb_loss_source():
    """Source lines of the inline bb loss, dedented, from the reference training script."""
    path = os.path.join(ref_shim.REF_ROOT, "train_Point2Cyl_without_sketch.py")
    with open(path) as f:
        lines = f.readlines()
    block = lines[285:307]  # 1-based 286..307: from 'cur_batch_size' to the double mean
    src = textwrap.dedent("".join(l.replace("\t", "    ") for l in block))
    assert "W_sorted, label = torch.sort" in src and "cross_entropy" in src, src
    return src


def loss_case(ref, name, B, N, K, seed, norm_eig):
    data = synthetic.s_cyl(B, N, K, seed)
    g = torch.Generator().manual_seed(seed + 7)
    X_raw = data["normals"] + 0.3 * torch.randn(B, N, 3, generator=g)
    # logits correlated with the gt labels so that matching is non-trivial but well separated
    W_raw = torch.randn(B, N, 2 * K, generator=g)
    perm = torch.stack([torch.randperm(K, generator=g) for _ in range(B)])
    col = torch.gather(perm, 1, data["inst"]) * 2 + data["bb"]
    W_raw.scatter_add_(2, col[:, :, None], torch.full((B, N, 1), 2.5))
    pcs, gt_normals = data["pcs"], data["normals"]
    gt_extrusion_instances, gt_bb_labels = data["inst"], data["bb"]
    X_raw.requires_grad_(True)     # the reference's own autograd gives the gradient goldens (dX_raw, dW_raw)
    W_raw.requires_grad_(True)

    X = F.normalize(X_raw, p=2, dim=2, eps=1e-12)
    W_2K = torch.softmax(W_raw, dim=2)
    W_barrel, W_barrel_bb = W_2K[:, :, ::2], W_raw[:, :, ::2]
    W_base, W_base_bb = W_2K[:, :, 1::2], W_raw[:, :, 1::2]
    W = W_barrel + W_base
    total, l_n, l_seg, matching_indices, mask = ref.losses.compute_all_losses(
        pcs, W, gt_extrusion_instances, X, gt_normals, 1.0, 1.0, return_match_indices=True)
    ns = dict(torch=torch, F=F, W=W, matching_indices=matching_indices, mask=mask, K=K, NUM_POINT=N,
              batch_size=B, sampled_pcs=pcs, W_barrel_bb=W_barrel_bb, W_base_bb=W_base_bb,
              gt_bb_labels=gt_bb_labels)
    exec(_inline_bb_loss_source(), ns)
    l_bb = ns["total_bb_loss"]
    mask_gt = ref.losses.get_mask_gt(gt_extrusion_instances, K)
    gi = matching_indices.unsqueeze(1).expand(B, N, K)
    E_AX = ref.data_utils.estimate_extrusion_axis(
        X, torch.gather(W_barrel, 2, gi), torch.gather(W_base, 2, gi), gt_bb_labels,
        gt_extrusion_instances, normalize=norm_eig)
    ext = ref.losses.compute_normal_loss(E_AX, data["axes"], angle_diff=False, collapse=False)
    l_ax = torch.mean(ref.losses.reduce_mean_masked_instance(ext, mask_gt))
    centers = ref.data_utils.estimate_extrusion_centers(torch.gather(W, 2, gi), pcs)
    cd = torch.square(centers - data["centers"]).sum(dim=-1)
    l_c = torch.mean(ref.losses.reduce_mean_masked_instance(cd, mask_gt))
    # total of the training script with all five multipliers = 1 (train_...:314,339,353)
    (total + l_bb + l_ax + l_c).backward()
    dX_raw, dW_raw = X_raw.grad.clone(), W_raw.grad.clone()
    X_raw, W_raw, X, W, W_barrel, W_base = (t.detach() for t in (X_raw, W_raw, X, W, W_barrel, W_base))
    total, l_n, l_seg, l_bb, l_ax, l_c, E_AX, centers = (t.detach() for t in (total, l_n, l_seg, l_bb, l_ax, l_c,
                                                                                E_AX, centers))
    hard = ref.losses.hard_W_encoding(W, to_null_mask=True)
    np.savez_compressed(
        os.path.join(HERE, name),
        X_raw=np32(X_raw), W_raw=np32(W_raw), dX_raw=np32(dX_raw), dW_raw=np32(dW_raw), total=np32(total), normal=np32(l_n), miou=np32(l_seg),
        bb=np32(l_bb), axis=np32(l_ax), center=np32(l_c), matching_indices=np32(matching_indices),
        mask=np32(ns["mask"] > 0), E_AX=np32(E_AX), centers=np32(centers),
        hard_argmax=np32(hard.argmax(-1)).astype(np.int8), hard_rowsum=np32(hard.sum(-1)).astype(np.int8),
        seg_iou=np32(ref.losses.compute_segmentation_iou(W, gt_extrusion_instances, matching_indices,
                                                         (ns["mask"] > 0).float())),
        normal_diff=np32(ref.losses.compute_normal_difference(X, gt_normals)),
        meta=np.array([B, N, K, seed, int(norm_eig)], dtype=np.int64))
    print("wrote", name)


def sample_grad(g, limit=2048):
    """Parameter gradients are stored whole when small, else as their first `limit` entries plus the L2 norm."""
    flat = g.detach().reshape(-1)
    return np32(flat[:limit]), np.float64(flat.double().norm())


def train_case(ref, name, B, N, K, seed, bn_eval=False):
    """One training step of the reference (train_Point2Cyl_without_sketch.py:244-367, all five multipliers 1):
    real backbone module in train mode, a fixed seeded dropout mask, the reference's loss functions and inline bb
    loss, torch autograd -> gradients of every parameter."""
    data = synthetic.s_cyl(B, N, K, seed)
    sd = orc.init_state_dict(output_sizes=(3, 2 * K), seed=seed)
    net = ref.net.backbone(output_sizes=[3, 2 * K])
    net.load_state_dict(sd, strict=True)
    net.train(not bn_eval)      # bn_eval: BatchNorm on running statistics (F.dropout stays active, :60) - the
    #                             well-conditioned variant of the same step, see tests/test_gpu_backward.py
    mask = (torch.rand(B, 128, N, generator=torch.Generator().manual_seed(seed + 3)) > 0.5).float() * 2.0
    real_dropout = ref.net.F.dropout
    ref.net.F.dropout = lambda x, p=0.5, **kw: x * mask
    try:
        torch.manual_seed(seed)
        s1 = torch.randint(0, N, (B,), dtype=torch.long)
        s2 = torch.randint(0, 512, (B,), dtype=torch.long)
        torch.manual_seed(seed)
        X_raw, W_raw = net(data["pcs"])
    finally:
        ref.net.F.dropout = real_dropout
    pcs, gt_normals, inst, bb = data["pcs"], data["normals"], data["inst"], data["bb"]
    X = F.normalize(X_raw, p=2, dim=2, eps=1e-12)
    W_2K = torch.softmax(W_raw, dim=2)
    W_barrel, W_barrel_bb = W_2K[:, :, ::2], W_raw[:, :, ::2]
    W_base, W_base_bb = W_2K[:, :, 1::2], W_raw[:, :, 1::2]
    W = W_barrel + W_base
    total, l_n, l_seg, matching_indices, mask_m = ref.losses.compute_all_losses(
        pcs, W, inst, X, gt_normals, 1.0, 1.0, return_match_indices=True)
    ns = dict(torch=torch, F=F, W=W, matching_indices=matching_indices, mask=mask_m, K=K, NUM_POINT=N,
              batch_size=B, sampled_pcs=pcs, W_barrel_bb=W_barrel_bb, W_base_bb=W_base_bb, gt_bb_labels=bb)
    exec(_inline_bb_loss_source(), ns)
    l_bb = ns["total_bb_loss"]
    mask_gt = ref.losses.get_mask_gt(inst, K)
    gi = matching_indices.unsqueeze(1).expand(B, N, K)
    E_AX = ref.data_utils.estimate_extrusion_axis(X, torch.gather(W_barrel, 2, gi), torch.gather(W_base, 2, gi), bb,
                                                  inst, normalize=False)
    ext = ref.losses.compute_normal_loss(E_AX, data["axes"], angle_diff=False, collapse=False)
    l_ax = torch.mean(ref.losses.reduce_mean_masked_instance(ext, mask_gt))
    centers = ref.data_utils.estimate_extrusion_centers(torch.gather(W, 2, gi), pcs)
    l_c = torch.mean(ref.losses.reduce_mean_masked_instance(torch.square(centers - data["centers"]).sum(dim=-1),
                                                            mask_gt))
    loss = total + l_bb + l_ax + l_c
    loss.backward()
    out = dict(loss=np32(loss), s1=np32(s1), s2=np32(s2), matching_indices=np32(matching_indices),
               X_raw=np32(X_raw), W_raw=np32(W_raw), meta=np.array([B, N, K, seed], dtype=np.int64))
    for k, p in net.named_parameters():
        out["grad_" + k], out["gnorm_" + k] = sample_grad(p.grad)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name)


def projection_case(ref, name, B, N, K, S, seed):
    """Second-wave closed forms run by the reference's own code (data_utils.py:1014-1417, :1650-1730); the CPU
    generator is re-seeded before each call so the consumer can reproduce the randint stream.  Cloud 0 loses one
    segment entirely except for a single point and one segment is relabelled base-only, so both `<= 1 member`
    branches (:1042, :1054) are exercised; the axes include an exact +z, a near +z and a -z axis."""
    data = synthetic.s_cyl(B, N, K, seed)
    P, X, inst, bb = data["pcs"], data["normals"], data["inst"].clone(), data["bb"].clone()
    axes, centers = data["axes"].clone(), data["centers"].clone()
    g = torch.Generator().manual_seed(seed + 11)
    # fill the zero-padded gt slots with random unit axes / centres so every (b,k) rotation is non-trivial
    pad = axes.norm(dim=-1) == 0
    rnd_ax = torch.nn.functional.normalize(torch.randn(B, K, 3, generator=g), dim=-1)
    axes[pad] = rnd_ax[pad]
    centers[pad] = (torch.rand(B, K, 3, generator=g) - 0.5)[pad]
    axes[0, 0] = torch.tensor([0.0, 0.0, 1.0])
    axes[1, 0] = torch.nn.functional.normalize(torch.tensor([1e-4, -2e-4, 1.0]), dim=-1)
    axes[min(2, B - 1), 1 % K] = torch.tensor([0.0, 0.0, -1.0])
    bb[0][inst[0] == 0] = 1                       # segment 0 of cloud 0: no barrel point at all
    first = (bb[0] == 1).nonzero()[0]
    bb[0, first] = 0
    inst[0, first] = 0                            # ... except exactly one (still "not found")
    out = {}
    torch.manual_seed(seed)
    Pp, Xp, sc = ref.data_utils.sketch_implicit_projection(P, X, inst, bb, axes, centers, num_points_to_sample=S)
    out.update(P_proj=np32(Pp), X_proj=np32(Xp), scales=np32(sc))
    torch.manual_seed(seed)
    Pp2, Xp2, sc2, found2 = ref.data_utils.sketch_implicit_projection2(P, X, inst, bb, axes, centers,
                                                                       num_points_to_sample=S)
    assert torch.equal(Pp2, Pp) and torch.equal(sc2, sc)
    out.update(found=np32(found2))
    Pp3, Xp3, sc3, found3 = ref.data_utils.sketch_implicit_projection3(P, X, inst, bb, axes, centers,
                                                                       num_points_to_sample=N)
    out.update(P_proj3=np32(Pp3), X_proj3=np32(Xp3), scales3=np32(sc3), found3=np32(found3))
    torch.manual_seed(seed + 1)
    ext, found_e = ref.data_utils.get_extrusion_extents(P, inst, bb, axes, centers, num_points_to_sample=S)
    out.update(extents=np32(ext), found_ext=np32(found_e))
    np.savez_compressed(os.path.join(HERE, name), inst=np32(inst).astype(np.int8), bb=np32(bb).astype(np.int8),
                        axes=np32(axes), centers=np32(centers), meta=np.array([B, N, K, S, seed], dtype=np.int64),
                        **out)
    print("wrote", name)


def _sketch_loss_source():
    """The implicit-sketch loss lines of the with-sketch trainer (train_Point2Cyl.py:608-672), dedented: from
    `if WITH_IM_LOSS:` to `im_loss += latent_loss`.  Not an importable function, so the reference's own lines are
    exec'd (same approach as the inline bb loss)."""
    path = os.path.join(ref_shim.REF_ROOT, "train_Point2Cyl.py")
    with open(path) as f:
        lines = f.readlines()
    block = lines[607:672]
    src = textwrap.dedent("".join(l.replace("\t", "    ") for l in block))
    assert src.startswith("if WITH_IM_LOSS:") and src.rstrip().endswith("im_loss += latent_loss"), src[:200]
    assert "nonmnfld_pnts = sampler.get_points(sk_pnts)" in src and "torch.min(values, dim=-1)" in src
    return src


def igr_inputs(B, K, S, seed):
    """Synthetic sketches: per (cloud, segment) S points on an ellipse with outward unit normals (the 'gt sketch'),
    and a jittered, rotated copy standing for the projected prediction.  -> gt_sketches (B,K,S,4), global_pc
    (B*K,S,4), mask_gt (B,K) bool (cloud b has 1 + (b % K) instances)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(B, K, S, generator=g) * (2 * np.pi)
    a = 0.3 + 0.6 * torch.rand(B, K, 1, generator=g)
    b = 0.3 + 0.6 * torch.rand(B, K, 1, generator=g)
    pts = torch.stack([a * torch.cos(t), b * torch.sin(t)], dim=-1)
    nrm = F.normalize(torch.stack([b * torch.cos(t), a * torch.sin(t)], dim=-1), dim=-1)
    gt_sketches = torch.cat([pts, nrm], dim=-1)
    ang = torch.rand(B, K, 1, 1, generator=g) * 0.3
    rot = torch.cat([torch.cat([torch.cos(ang), -torch.sin(ang)], -1), torch.cat([torch.sin(ang), torch.cos(ang)], -1)], -2)
    p2 = pts @ rot.transpose(-1, -2) + 0.01 * torch.randn(B, K, S, 2, generator=g)
    n2 = F.normalize(nrm @ rot.transpose(-1, -2) + 0.05 * torch.randn(B, K, S, 2, generator=g), dim=-1)
    global_pc = torch.cat([p2, n2], dim=-1).reshape(B * K, S, 4)
    n_inst = torch.tensor([1 + (i % K) for i in range(B)])
    mask_gt = torch.arange(K)[None, :] < n_inst[:, None]
    return gt_sketches, global_pc, mask_gt


def igr_case(ref, name, B, K, S, seed, is_l2):
    """latent codes, the implicit network's prediction and input gradient, the four loss terms
    and every parameter gradient, by the reference's own modules (IGR/network.py, IGR/sampler.py) and loss lines."""
    net_ref = ref_shim.load_igr()
    nw, sm = net_ref.network, net_ref.sampler
    gt_sketches, global_pc, mask_gt = igr_inputs(B, K, S, seed)
    implicit_net = nw.ImplicitNet(d_in=igr.D_IN + igr.LATENT_SIZE, dims=list(igr.IMPLICIT_DIMS), skip_in=list(igr.IMPLICIT_SKIP),
                                  geometric_init=True, radius_init=1, beta=100)
    implicit_net.load_state_dict(igr.implicit_init(seed=seed), strict=True)
    pn_encoder = nw.PointNetEncoder(igr.LATENT_SIZE, igr.D_IN, with_normals=True)
    pn_encoder.load_state_dict(igr.encoder_init(seed=seed + 1), strict=True)
    loaded_pn_encoder = nw.PointNetEncoder(igr.LATENT_SIZE, igr.D_IN, with_normals=True)
    loaded_pn_encoder.load_state_dict(igr.encoder_init(seed=seed + 2), strict=True)
    pn_encoder.train()
    loaded_pn_encoder.train()
    # train_Point2Cyl.py:598-605
    batch_size, NUM_SK_POINT = B, S
    latent_codes = pn_encoder(global_pc)
    sk_pnts = gt_sketches[:, :, :, :2].view(batch_size * K, NUM_SK_POINT, 2)
    sk_normals = gt_sketches[:, :, :, -2:].view(batch_size * K, NUM_SK_POINT, 2)
    global_pc_gt = torch.cat((sk_pnts, sk_normals), dim=-1)
    latent_codes_gt = loaded_pn_encoder(global_pc_gt)
    latent_codes.retain_grad()
    ns = dict(torch=torch, sampler=sm.NormalPerPoint(1.8, 0.01), add_latent=nw.add_latent, gradient=nw.gradient,
              implicit_net=implicit_net, reduce_mean_masked_instance=ref_shim.load().losses.reduce_mean_masked_instance,
              mask_gt=mask_gt, sk_pnts=sk_pnts, sk_normals=sk_normals, latent_codes=latent_codes,
              latent_codes_gt=latent_codes_gt, batch_size=batch_size, K=K, WITH_IM_LOSS=True, IS_L2=is_l2,
              pcs=global_pc)
    torch.manual_seed(seed + 3)                      # the sampler draws randn_like then rand from the global generator
    exec(_sketch_loss_source(), ns)
    ns["im_loss"].backward()
    out = dict(meta=np.array([B, K, S, seed, int(is_l2)]), gt_sketches=np32(gt_sketches), global_pc=np32(global_pc),
               mask_gt=mask_gt.numpy(), latent=np32(latent_codes), latent_gt=np32(latent_codes_gt),
               d_latent=np32(latent_codes.grad),
               nonmnfld_pnts=np32(ns["nonmnfld_pnts"][:, -2:]), sk_pred=np32(ns["sk_pred"].reshape(-1, 1)),
               mnfld_grad=np32(ns["mnfld_grad"].reshape(-1, 2)), nonmnfld_grad=np32(ns["nonmnfld_grad"].reshape(-1, 2)))
    for k in ("im_loss", "mnfld_loss", "grad_loss", "normals_loss", "latent_loss"):
        out[k] = np32(ns[k])
    for prefix, mod in (("net", implicit_net), ("enc", pn_encoder), ("encgt", loaded_pn_encoder)):
        for pname, p in mod.named_parameters():
            smp, nrm_ = sample_grad(p.grad)
            out[f"grad_{prefix}.{pname}"] = smp
            out[f"gradnorm_{prefix}.{pname}"] = nrm_
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: float(out[k]) for k in ("im_loss", "mnfld_loss", "grad_loss", "normals_loss", "latent_loss")})


def main():
    assert ref_shim.available(), "reference tree not present"
    torch.set_num_threads(8)
    ref = ref_shim.load()
    cyl = synthetic.s_cyl(2, 1024, 4, seed=0)["pcs"]
    pointops_case(ref, "pointops_cyl_n1024.npz", cyl, npoint=128, radius=0.2, nsample=32, seed=0)
    uni = synthetic.s_uniform(2, 2048, seed=1)
    pointops_case(ref, "pointops_uniform_n2048.npz", uni, npoint=256, radius=0.2, nsample=64, seed=1)
    backbone_case(ref, "backbone_b2_n1024_k4.npz", 2, 1024, 4, seed=0)
    backbone_case(ref, "backbone_b1_n1024_k4.npz", 1, 1024, 4, seed=3)   # BASELINE.json config 1
    loss_case(ref, "loss_b2_n1024_k4.npz", 2, 1024, 4, seed=0, norm_eig=False)
    loss_case(ref, "loss_b3_n2048_k8_normeig.npz", 3, 2048, 8, seed=5, norm_eig=True)
    projection_case(ref, "projection_b3_n512_k4.npz", 3, 512, 4, 128, seed=2)
    train_case(ref, "train_b2_n1024_k4.npz", 2, 1024, 4, seed=0)
    train_case(ref, "train_bneval_b2_n1024_k4.npz", 2, 1024, 4, seed=0, bn_eval=True)
    igr_case(ref, "igr_b3_k2_s64.npz", 3, 2, 64, seed=0, is_l2=False)
    igr_case(ref, "igr_b2_k4_s128_l2.npz", 2, 4, 128, seed=4, is_l2=True)


if __name__ == "__main__":
    main()
