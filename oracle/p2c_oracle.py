"""CPU oracle for the Point2Cyl forward+loss hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU restatement of the reference algorithm (the
reference itself is a PyTorch program, so its arithmetic is torch's).  It is the
checker for the CUDA path; it is never the thing shipped or measured:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product package
(point2cyl_b200/) never does.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the pin is the reference run in the authoring container:
tests/golden/make_golden.py imports the real reference through
oracle/ref_shim.py, runs both on the same seeded inputs and commits the
reference outputs under tests/golden/; tests/test_oracle_golden.py checks this
restatement against those files (indices bit-exact, floats to 1e-6).
One exception, PARITY UNPINNED: torchgeometry==0.1.2's angle_axis_to_rotation_matrix
(used only by the projection functions) is absent offline and restated
from its published algorithm; see the note above `angle_axis_to_rotation_matrix` below.

Every function cites the reference file:line it follows (paths relative to the
upstream repo root).  Layout conventions are the reference's: clouds are
(B, N, 3) point-major at function level and (B, C, N) channel-first at
nn.Module level; indices are int64.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------
# L2: point-set ops  (models/pointnet_util.py)
# ----------------------------------------------------------------------------------------------


def square_distance(src: Tensor, dst: Tensor) -> Tensor:
    """Expanded-form pairwise squared distance, models/pointnet_util.py:19-40.

    d[b,i,j] = ((-2 * <src_i, dst_j>) + |src_i|^2) + |dst_j|^2 with the dot product from a
    batched matmul; the association order matters for the bit-exact ball-query membership test.
    """
    B, S, _ = src.shape
    M = dst.shape[1]
    d = torch.matmul(src, dst.transpose(1, 2)) * -2
    d = d + (src ** 2).sum(-1).reshape(B, S, 1)
    d = d + (dst ** 2).sum(-1).reshape(B, 1, M)
    return d


def square_distance_fma_emulated(src: Tensor, dst: Tensor) -> Tensor:
    """The same quantity with the K=3 dot product spelled out as the FMA chain
    fma(a2,b2, fma(a1,b1, a0*b0)) (SURVEY.md section 7 'Bit-exact ball query').

    float64 products of float32 operands are exact, so round32(a*b + c) emulates a float32
    FMA up to (vanishingly rare) double rounding.  Used to pin the recipe the CUDA kernels
    implement against what torch.matmul does on this CPU.
    """
    a = src.double()
    b = dst.double()

    def r32(x):
        return x.float().double()

    t = r32(a[:, :, None, 0] * b[:, None, :, 0])
    t = r32(a[:, :, None, 1] * b[:, None, :, 1] + t)
    t = r32(a[:, :, None, 2] * b[:, None, :, 2] + t)
    dot = t.float()
    d = dot * -2
    d = d + (src ** 2).sum(-1)[:, :, None]
    d = d + (dst ** 2).sum(-1)[:, None, :]
    return d


def gather_points(points: Tensor, idx: Tensor) -> Tensor:
    """points[b, idx[b, ...], :]  — models/pointnet_util.py:43-60 (index_points)."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.gather(points, 1, flat[:, :, None].expand(-1, -1, points.shape[-1]))
    return out.reshape(*idx.shape, points.shape[-1])


def draw_fps_start(B: int, N: int) -> Tensor:
    """The reference draws the first centroid from the CPU generator, pointnet_util.py:75."""
    return torch.randint(0, N, (B,), dtype=torch.long)


def farthest_point_sample(xyz: Tensor, npoint: int, start: Optional[Tensor] = None) -> Tensor:
    """Iterative farthest point sampling, models/pointnet_util.py:63-84.

    Running distance starts at 1e10 (:74); each round records the current farthest point (:78),
    computes sum((xyz - c)^2, -1) (:80), lowers the running distance where strictly smaller
    (:81-82) and picks the first arg-max (:83).
    """
    B, N, _ = xyz.shape
    if start is None:
        start = draw_fps_start(B, N)
    far = start.to(torch.long).clone()
    running = torch.full((B, N), 1e10, dtype=xyz.dtype)
    picked = torch.empty(B, npoint, dtype=torch.long)
    rows = torch.arange(B)
    for i in range(npoint):
        picked[:, i] = far
        c = xyz[rows, far].reshape(B, 1, 3)
        d = ((xyz - c) ** 2).sum(-1)
        running = torch.where(d < running, d, running)
        far = running.max(dim=-1).indices
    return picked


def query_ball_point(radius: float, nsample: int, xyz: Tensor, new_xyz: Tensor) -> Tensor:
    """Ball query, models/pointnet_util.py:87-107.

    Reference: label every out-of-ball point N, sort, keep the first nsample, replace the N's
    by the first kept index.  Equivalent statement used here: the first nsample in-ball indices
    in ascending order, padded with the first hit.  "Out of ball" is d > float32(radius**2)
    (:102; the Python double is cast to the tensor dtype by the comparison).
    """
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    d = square_distance(new_xyz, xyz)
    r2 = torch.tensor(radius ** 2, dtype=d.dtype)
    inside = ~(d > r2)
    # rank of each in-ball point among the in-ball points of its row
    rank = inside.long().cumsum(-1) - 1
    keep = inside & (rank < nsample)
    out = torch.full((B, S, nsample), N, dtype=torch.long)
    b, s, n = keep.nonzero(as_tuple=True)
    out[b, s, rank[b, s, n]] = n
    first = out[:, :, :1].expand(-1, -1, nsample)
    return torch.where(out == N, first, out)


def sample_and_group(npoint: int, radius: float, nsample: int, xyz: Tensor,
                     points: Optional[Tensor], start: Optional[Tensor] = None, forced=None):
    """FPS -> centres -> ball query -> centred neighbourhoods [+ features],
    models/pointnet_util.py:110-143.  Returns (new_xyz, new_points, fps_idx, group_idx).
    forced = (fps_idx, group_idx): use these index tensors instead of computing them (float64 adjudication runs
    take the float32 run's discrete choices so that only the arithmetic differs)."""
    B = xyz.shape[0]
    if forced is not None:
        fps_idx, group_idx = forced
        new_xyz = gather_points(xyz, fps_idx)
    else:
        fps_idx = farthest_point_sample(xyz, npoint, start)
        new_xyz = gather_points(xyz, fps_idx)
        group_idx = query_ball_point(radius, nsample, xyz, new_xyz)
    local = gather_points(xyz, group_idx) - new_xyz.reshape(B, npoint, 1, 3)
    if points is not None:
        local = torch.cat([local, gather_points(points, group_idx)], dim=-1)  # xyz first (:137)
    return new_xyz, local, fps_idx, group_idx


def sample_and_group_all(xyz: Tensor, points: Optional[Tensor]):
    """One group holding every point, centre at the origin, models/pointnet_util.py:146-163."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, dtype=xyz.dtype)
    grouped = xyz.reshape(B, 1, N, C)
    if points is not None:
        grouped = torch.cat([grouped, points.reshape(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped


def three_nn_interpolate(xyz1: Tensor, xyz2: Tensor, points2: Tensor, forced=None):
    """Inverse-distance 3-NN interpolation, models/pointnet_util.py:298-308.

    Uses the expanded-form distance (it can be slightly negative; the reference does not clamp)
    and a full sort whose first three entries are the neighbours.  Returns (interp, idx, weight).
    forced = (idx, weight): take the neighbours and weights of another (float32) run.
    """
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        return points2.repeat(1, N, 1), None, None
    if forced is not None:
        order, w = forced[0], forced[1].to(points2.dtype)
    else:
        d, order = square_distance(xyz1, xyz2).sort(dim=-1)
        d, order = d[:, :, :3], order[:, :, :3]
        recip = 1.0 / (d + 1e-8)
        w = recip / recip.sum(dim=2, keepdim=True)
    interp = (gather_points(points2, order) * w.reshape(B, N, 3, 1)).sum(dim=2)
    return interp, order, w


# ----------------------------------------------------------------------------------------------
# L3: backbone, functional over a state_dict  (models/pointnet_extrusion.py)
# ----------------------------------------------------------------------------------------------

SA_SPECS = (  # models/pointnet_extrusion.py:21-23
    dict(name="sa1", npoint=512, radius=0.2, nsample=64, mlp=(64, 64, 128), group_all=False),
    dict(name="sa2", npoint=128, radius=0.4, nsample=64, mlp=(128, 128, 256), group_all=False),
    dict(name="sa3", npoint=None, radius=None, nsample=None, mlp=(256, 512, 1024), group_all=True),
)
FP_SPECS = (  # models/pointnet_extrusion.py:25-27
    dict(name="fp3", mlp=(256, 256)),
    dict(name="fp2", mlp=(256, 128)),
    dict(name="fp1", mlp=(128, 128, 128)),
)


def init_state_dict(output_sizes: Sequence[int] = (3, 16), seed: int = 0,
                    normal_channel: bool = False) -> Dict[str, Tensor]:
    """Random-init weights with the reference's state_dict keys and shapes.

    Kaiming-uniform-like ranges as nn.Conv default; BN affine perturbed away from (1, 0) and
    running stats away from (0, 1) so that eval-mode tests are not trivially the identity.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    extra = 3 if normal_channel else 0

    def conv(prefix, cin, cout, two_d):
        bound = 1.0 / math.sqrt(cin)
        shape = (cout, cin, 1, 1) if two_d else (cout, cin, 1)
        sd[prefix + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[prefix + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(prefix, c):
        sd[prefix + ".weight"] = 1.0 + 0.2 * (torch.rand(c, generator=g) - 0.5)
        sd[prefix + ".bias"] = 0.2 * (torch.rand(c, generator=g) - 0.5)
        sd[prefix + ".running_mean"] = 0.1 * (torch.rand(c, generator=g) - 0.5)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    cin_sa = (3 + extra, 128 + 3, 256 + 3)
    for spec, cin in zip(SA_SPECS, cin_sa):
        for i, cout in enumerate(spec["mlp"]):
            conv(f"{spec['name']}.mlp_convs.{i}", cin, cout, True)
            bn(f"{spec['name']}.mlp_bns.{i}", cout)
            cin = cout
    cin_fp = (1024 + 256, 256 + 128, 128 + extra)
    for spec, cin in zip(FP_SPECS, cin_fp):
        for i, cout in enumerate(spec["mlp"]):
            conv(f"{spec['name']}.mlp_convs.{i}", cin, cout, False)
            bn(f"{spec['name']}.mlp_bns.{i}", cout)
            cin = cout
    conv("fc1", 128, 128, False)
    bn("bn1", 128)
    for j, o in enumerate(output_sizes):
        conv(f"fc2.{j}", 128, o, False)
    return sd


def _bn(x: Tensor, sd: Dict[str, Tensor], prefix: str, training: bool, momentum: float,
        new_stats: Optional[dict]) -> Tensor:
    """Stock BatchNorm (eps 1e-5): batch statistics in training, running statistics in eval."""
    rm = sd[prefix + ".running_mean"].clone()
    rv = sd[prefix + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"],
                     training=training, momentum=momentum, eps=1e-5)
    if new_stats is not None and training:
        new_stats[prefix + ".running_mean"] = rm
        new_stats[prefix + ".running_var"] = rv
    return y


def set_abstraction(sd, spec, xyz_cf: Tensor, feats_cf: Optional[Tensor], training: bool,
                    momentum: float = 0.1, start: Optional[Tensor] = None, new_stats=None,
                    trace: Optional[dict] = None, forced: Optional[dict] = None):
    """PointNetSetAbstraction.forward, models/pointnet_util.py:181-207 (channel-first I/O)."""
    name = spec["name"]
    xyz = xyz_cf.permute(0, 2, 1)
    feats = feats_cf.permute(0, 2, 1) if feats_cf is not None else None
    if spec["group_all"]:
        new_xyz, grouped = sample_and_group_all(xyz, feats)
    else:
        f = None if forced is None else (forced[name + ".fps_idx"], forced[name + ".group_idx"])
        new_xyz, grouped, fps_idx, group_idx = sample_and_group(
            spec["npoint"], spec["radius"], spec["nsample"], xyz, feats, start, forced=f)
        if trace is not None:
            trace[name + ".fps_idx"] = fps_idx
            trace[name + ".group_idx"] = group_idx
    h = grouped.permute(0, 3, 2, 1)  # (B, C, nsample, S)  (:200)
    for i in range(len(spec["mlp"])):
        h = F.conv2d(h, sd[f"{name}.mlp_convs.{i}.weight"], sd[f"{name}.mlp_convs.{i}.bias"])
        h = F.relu(_bn(h, sd, f"{name}.mlp_bns.{i}", training, momentum, new_stats))
    pooled = h.max(dim=2).values  # (:205)
    return new_xyz.permute(0, 2, 1), pooled


def feature_propagation(sd, spec, xyz1_cf, xyz2_cf, points1_cf, points2_cf, training: bool,
                        momentum: float = 0.1, new_stats=None, trace: Optional[dict] = None,
                        forced: Optional[dict] = None):
    """PointNetFeaturePropagation.forward, models/pointnet_util.py:283-320."""
    name = spec["name"]
    xyz1 = xyz1_cf.permute(0, 2, 1)
    xyz2 = xyz2_cf.permute(0, 2, 1)
    f = None
    if forced is not None and xyz2.shape[1] > 1:
        f = (forced[name + ".nn_idx"], forced[name + ".nn_w"])
    interp, nn_idx, nn_w = three_nn_interpolate(xyz1, xyz2, points2_cf.permute(0, 2, 1), forced=f)
    if trace is not None and nn_idx is not None:
        trace[name + ".nn_idx"] = nn_idx
        trace[name + ".nn_w"] = nn_w
    if points1_cf is not None:
        interp = torch.cat([points1_cf.permute(0, 2, 1), interp], dim=-1)  # skip feats first (:312)
    h = interp.permute(0, 2, 1)
    for i in range(len(spec["mlp"])):
        h = F.conv1d(h, sd[f"{name}.mlp_convs.{i}.weight"], sd[f"{name}.mlp_convs.{i}.bias"])
        h = F.relu(_bn(h, sd, f"{name}.mlp_bns.{i}", training, momentum, new_stats))
    return h


def backbone_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = True,
                     momentum: float = 0.1, fps_start: Optional[Sequence[Tensor]] = None,
                     dropout_mask: Optional[Tensor] = None, new_stats: Optional[dict] = None,
                     trace: Optional[dict] = None, forced: Optional[dict] = None) -> List[Tensor]:
    """backbone.forward, models/pointnet_extrusion.py:37-66.

    fps_start: (start_sa1 (B,), start_sa2 (B,)) or None to draw them like the reference does
    (one CPU randint per level, sa1 first).  dropout_mask: (B,128,N) multiplicative mask that
    already contains the 1/(1-p) scale (what F.dropout(ones) returns), or None for identity —
    the reference applies F.dropout(p=0.5) unconditionally (:60).
    forced: the `trace` of another run - its FPS / ball-query / 3-NN indices and 3-NN weights are used instead of
    being recomputed.  With sd and x in float64 this gives the exact-arithmetic value of the same discrete
    computation (tests: float64 adjudication of the float32 tolerances).
    """
    xc = x.transpose(2, 1)
    pos = xc[:, :3, :]
    feats = xc[:, 3:, :] if xc.shape[1] > 3 else None
    s1 = fps_start[0] if fps_start is not None else None
    s2 = fps_start[1] if fps_start is not None else None
    l1_xyz, l1 = set_abstraction(sd, SA_SPECS[0], pos, feats, training, momentum, s1, new_stats, trace, forced)
    l2_xyz, l2 = set_abstraction(sd, SA_SPECS[1], l1_xyz, l1, training, momentum, s2, new_stats, trace, forced)
    l3_xyz, l3 = set_abstraction(sd, SA_SPECS[2], l2_xyz, l2, training, momentum, None, new_stats, trace, forced)
    l4 = feature_propagation(sd, FP_SPECS[0], l2_xyz, l3_xyz, l2, l3, training, momentum, new_stats, trace, forced)
    l5 = feature_propagation(sd, FP_SPECS[1], l1_xyz, l2_xyz, l1, l4, training, momentum, new_stats, trace, forced)
    l6 = feature_propagation(sd, FP_SPECS[2], pos, l1_xyz, feats, l5, training, momentum, new_stats, trace, forced)
    h = F.conv1d(l6, sd["fc1.weight"], sd["fc1.bias"])
    h = F.relu(_bn(h, sd, "bn1", training, momentum, new_stats))
    if dropout_mask is not None:
        h = h * dropout_mask
    if trace is not None:
        trace.update(l1_xyz=l1_xyz, l1=l1, l2_xyz=l2_xyz, l2=l2, l3=l3, l4=l4, l5=l5, l6=l6, head=h)
    outs = []
    j = 0
    while f"fc2.{j}.weight" in sd:
        outs.append(F.conv1d(h, sd[f"fc2.{j}.weight"], sd[f"fc2.{j}.bias"]).transpose(1, 2))
        j += 1
    return outs


# ----------------------------------------------------------------------------------------------
# L4: losses and closed-form fitting  (losses.py, data_utils.py, train_Point2Cyl_without_sketch.py)
# ----------------------------------------------------------------------------------------------


def sequence_mask(lengths: Tensor, maxlen: Optional[int] = None) -> Tensor:
    """losses.py:70-76."""
    if maxlen is None:
        maxlen = int(lengths.max())
    return torch.arange(maxlen)[None, :] < lengths[:, None]


def get_mask_gt(I_gt: Tensor, n_max_instances: int) -> Tensor:
    """losses.py:78-81: instance k exists iff k <= max label."""
    return sequence_mask(I_gt.max(dim=1).values + 1, n_max_instances)


def reduce_mean_masked_instance(loss: Tensor, mask_gt: Tensor) -> Tensor:
    """losses.py:83-88: per-sample mean over existing instances (0 where there are none)."""
    kept = torch.where(mask_gt, loss, torch.zeros_like(loss)).sum(dim=1)
    cnt = mask_gt.to(loss.dtype).sum(dim=1)
    return torch.where(cnt > 0, kept / cnt, torch.zeros_like(kept))


def hungarian_matching(W_pred: Tensor, I_gt: Tensor):
    """losses.py:22-52.  Per sample: relaxed-IoU cost between gt one-hot columns and predicted
    columns, maximised with scipy's linear_sum_assignment; slots >= n_gt stay 0.
    Returns (matching_indices (B,K) int64, mask (B,K) bool)."""
    B, N, K = W_pred.shape
    match = torch.zeros(B, K, dtype=torch.long)
    mask = torch.zeros(B, K, dtype=torch.bool)
    for b in range(B):
        n_gt = int(I_gt[b].max()) + 1
        onehot = torch.eye(n_gt + 1, dtype=W_pred.dtype)[I_gt[b]]   # label -1 -> last row (:38)
        inter = onehot.t() @ W_pred[b]
        union = onehot.sum(0)[:, None] + W_pred[b].sum(0)[None, :] - inter
        score = (inter / union.clamp(min=1e-10))[:n_gt]
        _, cols = linear_sum_assignment(-score.detach().cpu().numpy())   # .cpu(): as losses.py:43
        match[b, :n_gt] = torch.from_numpy(cols).long()
        mask[b, :n_gt] = True
    return match, mask


def compute_miou_loss(W: Tensor, I_gt: Tensor, matching_indices: Tensor, div_eps: float = 1e-10):
    """losses.py:90-103: 1 - relaxed IoU per gt slot after re-ordering W's columns by the match."""
    B, N, K = W.shape
    L = matching_indices.shape[1]
    Wr = torch.gather(W, 2, matching_indices[:, None, :].expand(B, N, L))
    onehot = torch.eye(L + 2, dtype=W.dtype)[I_gt][:, :, :L]
    inter = (onehot * Wr).sum(dim=1)
    union = onehot.sum(dim=1) + Wr.sum(dim=1) - inter
    return 1.0 - inter / (union + div_eps), 1 - inter / N, Wr


def compute_normal_loss(normal: Tensor, normal_gt: Tensor, collapse: bool = True) -> Tensor:
    """losses.py:127-143 with angle_diff=False: 1 - |<n, n_gt>|, optionally averaged over points."""
    v = 1.0 - (normal * normal_gt).sum(dim=2).abs()
    return v.mean(dim=1) if collapse else v


def compute_all_losses(P, W, I_gt, X, X_gt, normal_loss_multiplier, miou_loss_multiplier):
    """losses.py:317-351 (collapse=True, return_match_indices=True)."""
    B, _, K = W.shape
    mask_gt = get_mask_gt(I_gt, K)
    normal_loss = compute_normal_loss(X, X_gt) if normal_loss_multiplier > 0 else torch.zeros(B, K)
    match, mask = hungarian_matching(W, I_gt)
    miou, _, _ = compute_miou_loss(W, I_gt, match)
    avg_miou = reduce_mean_masked_instance(miou, mask_gt)
    total_miou = avg_miou.mean()
    total_normal = normal_loss.mean()
    total = miou_loss_multiplier * total_miou + normal_loss_multiplier * total_normal
    return total, total_normal, total_miou, match, mask


def bb_loss(W: Tensor, W_raw_barrel: Tensor, W_raw_base: Tensor, gt_bb: Tensor,
            match: Tensor, mask: Tensor) -> Tensor:
    """Base/barrel loss written inline in train_Point2Cyl_without_sketch.py:283-313.

    Columns of W re-ordered by the match, unmatched slots zeroed, softmax over K, ascending sort;
    the sort permutation indexes the *un-reordered* raw barrel/base logits (reference quirk,
    SURVEY.md a14); 2-way cross-entropy per (point, slot) weighted by the sorted probability.
    """
    B, N, K = W.shape
    Wr = torch.gather(W, 2, match[:, None, :].expand(B, N, K))
    Wr = torch.where(mask[:, None, :].expand(B, N, K), Wr, torch.zeros_like(Wr))
    Z, lab = torch.softmax(Wr, dim=-1).sort(dim=-1)
    logits = torch.stack([torch.gather(W_raw_barrel, 2, lab), torch.gather(W_raw_base, 2, lab)], -1)
    ce = F.cross_entropy(logits.reshape(B * N * K, 2), gt_bb[:, :, None].expand(B, N, K).reshape(-1),
                         reduction="none").reshape(B, N, K)
    return (ce * Z).sum(-1).mean(-1).mean()


def axis_scatter_matrices(X, W_barrel, W_base, gt_bb=None, gt_inst=None, normalize=False) -> Tensor:
    """Closed form of BtB - CtC in data_utils.py:118-163:
    M[b,k] = sum_n (wbar[n,k]^2 - wbase[n,k]^2) x_n x_n^T, with wbar/(sqrt(#gt barrel_k)+1) and
    wbase/(sqrt(#gt base_k)+1) when normalize (:133-160).  float64 accumulation."""
    B, N, K = W_barrel.shape
    wb, wc = W_barrel.double(), W_base.double()
    if normalize:
        inst = torch.stack([(gt_inst == k) for k in range(K)], -1)
        nb = (inst & (gt_bb == 0)[:, :, None]).float().sum(1).sqrt() + 1.0
        nc = (inst & (gt_bb == 1)[:, :, None]).float().sum(1).sqrt() + 1.0
        wb = wb / nb[:, None, :].double()
        wc = wc / nc[:, None, :].double()
    coef = wb ** 2 - wc ** 2
    Xd = X.double()
    return torch.einsum("bnk,bni,bnj->bkij", coef, Xd, Xd)


def estimate_extrusion_axis(X, W_barrel, W_base, gt_bb=None, gt_inst=None, normalize=False,
                            dense: bool = False) -> Tensor:
    """data_utils.py:99-177: eigenvector of the smallest eigenvalue of BtB - CtC per (b, k).

    dense=True walks the reference's own route (diag_embed to (B,N,N), four bmm's, fp32) and is
    what the CPU baseline times; dense=False uses the algebraically identical 3x3 scatter
    (float64) so large N is tractable in tests.  Eigenvector sign is arbitrary: compare by |dot|.
    """
    B, N, K = W_barrel.shape
    if not dense:
        M = axis_scatter_matrices(X, W_barrel, W_base, gt_bb, gt_inst, normalize)
        _, v = torch.linalg.eigh(M, UPLO="U")
        return v[..., 0].to(X.dtype)
    out = torch.zeros(B, K, 3)
    for k in range(K):
        Db = torch.diag_embed(W_barrel[:, :, k])
        Dc = torch.diag_embed(W_base[:, :, k])
        Bm = torch.bmm(Db, X)
        Cm = torch.bmm(Dc, X)
        if normalize:
            sel = (gt_inst == k)
            nb = (sel & (gt_bb == 0)).float().sum(-1).sqrt()
            nc = (sel & (gt_bb == 1)).float().sum(-1).sqrt()
            Bm = Bm / (nb[:, None, None] + 1.0)
            Cm = Cm / (nc[:, None, None] + 1.0)
        M = torch.bmm(Bm.transpose(1, 2), Bm) - torch.bmm(Cm.transpose(1, 2), Cm)
        _, v = torch.linalg.eigh(M, UPLO="U")   # torch.symeig(upper=True) of torch 1.8 (:170)
        out[:, k, :] = v[:, :, 0]
    return out


def estimate_extrusion_centers(W: Tensor, pcs: Tensor) -> Tensor:
    """data_utils.py:253-266: c[b,k] = mean_n W[b,n,k] * p[b,n]  (plain mean, not / sum W)."""
    return torch.einsum("bnk,bnc->bkc", W, pcs) / W.shape[1]


def postnet_split(X_raw: Tensor, W_raw: Tensor):
    """train_Point2Cyl_without_sketch.py:246-265: unit normals, softmax over 2K, even columns =
    barrel, odd = base, W = barrel + base."""
    X = F.normalize(X_raw, p=2, dim=2, eps=1e-12)
    W2K = torch.softmax(W_raw, dim=2)
    Wb, Wc = W2K[:, :, 0::2], W2K[:, :, 1::2]
    return X, Wb, Wc, Wb + Wc


def loss_block(pcs, X_raw, W_raw, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers,
               weights=(1.0, 1.0, 1.0, 1.0, 1.0), norm_eig=False, dense_axis=False) -> Dict[str, Tensor]:
    """The loss half of a training step, train_Point2Cyl_without_sketch.py:246-353, with all five
    --pred_* branches enabled.  weights = (seg, normal, bb, extrusion, centre)."""
    w_seg, w_n, w_bb, w_ext, w_c = weights
    K = W_raw.shape[2] // 2
    X, Wb, Wc, W = postnet_split(X_raw, W_raw)
    total, l_n, l_seg, match, mask = compute_all_losses(pcs, W, gt_inst, X, gt_normals, w_n, w_seg)
    l_bb = bb_loss(W, W_raw[:, :, 0::2], W_raw[:, :, 1::2], gt_bb, match, mask)
    total = total + w_bb * l_bb
    mask_gt = get_mask_gt(gt_inst, K)
    B, N, _ = W.shape
    gidx = match[:, None, :].expand(B, N, K)
    E_AX = estimate_extrusion_axis(X, torch.gather(Wb, 2, gidx), torch.gather(Wc, 2, gidx),
                                   gt_bb, gt_inst, normalize=norm_eig, dense=dense_axis)
    l_ax = reduce_mean_masked_instance(compute_normal_loss(E_AX, gt_axes, collapse=False), mask_gt).mean()
    total = total + w_ext * l_ax
    centers = estimate_extrusion_centers(torch.gather(W, 2, gidx), pcs)
    l_c = reduce_mean_masked_instance(((centers - gt_centers) ** 2).sum(-1), mask_gt).mean()
    total = total + w_c * l_c
    return dict(total=total, normal=l_n, miou=l_seg, bb=l_bb, axis=l_ax, center=l_c,
                matching_indices=match, mask=mask, E_AX=E_AX, centers=centers)


def forward_loss(sd, batch: Dict[str, Tensor], training=True, momentum=0.1, fps_start=None,
                 dropout_mask=None, weights=(1.0,) * 5, norm_eig=False, dense_axis=False,
                 trace=None, forced=None, new_stats=None) -> Dict[str, Tensor]:
    """One forward+loss pass (the unit BASELINE.json's metric counts clouds over)."""
    X_raw, W_raw = backbone_forward(sd, batch["pcs"], training, momentum, fps_start, dropout_mask,
                                    new_stats=new_stats, trace=trace, forced=forced)
    out = loss_block(batch["pcs"], X_raw, W_raw, batch["normals"], batch["inst"], batch["bb"],
                     batch["axes"], batch["centers"], weights, norm_eig, dense_axis)
    out.update(X_raw=X_raw, W_raw=W_raw)
    return out


# eval-side helpers -------------------------------------------------------------


def hard_W_encoding(W: Tensor, to_null_mask=False, W_null_threshold=0.005) -> Tensor:
    """losses.py:55-68: arg-max one-hot; columns with sum W < threshold*N zeroed when asked."""
    B, N, K = W.shape
    hard = torch.eye(K)[W.argmax(dim=2)].float()
    if to_null_mask:
        null = (W.sum(dim=1) < float(N) * W_null_threshold).float()
        hard = hard * (1.0 - null[:, None, :])
    return hard


def acos_safe(x: Tensor) -> Tensor:
    """losses.py:123-124."""
    return torch.acos(x.clamp(min=-1.0 + 1e-6, max=1.0 - 1e-6))


def compute_segmentation_iou(W, I_gt, matching_indices, mask) -> Tensor:
    """losses.py:106-109."""
    miou = 1 - compute_miou_loss(W, I_gt, matching_indices)[0]
    return (mask * miou).sum(dim=1) / mask.sum(dim=1)


def compute_normal_difference(X, X_gt, in_radians=True, collapse=True) -> Tensor:
    """losses.py:146-159."""
    d = acos_safe((X * X_gt).sum(dim=2).abs())
    if not in_radians:
        d = d * 180.0 / math.pi
    return d.mean(dim=1) if collapse else d


# projection / scale / extent closed forms -------------------------------------------------------
#
# PARITY UNPINNED for the rotation: the reference calls torchgeometry==0.1.2 `angle_axis_to_rotation_matrix`
# (requirements.txt:15), which is neither under /root/reference nor installable offline.  The function below
# restates its published algorithm (ceres-style Rodrigues with w = aa / (theta + 1e-6) for theta^2 > 1e-6, the
# first-order Taylor form otherwise, returned as a 4x4 homogeneous matrix).  oracle/ref_shim.py installs it as the
# `torchgeometry` stub so that the reference's OWN projection functions (data_utils.py:1014-1417, :1650-1730) can
# be run here to produce tests/golden/projection_*.npz; everything around the rotation is therefore pinned by the
# reference's code, the rotation itself only by this restatement (cross-checked against scipy's Rodrigues
# implementation in tests/test_oracle_golden.py - an independent check of the formula, not of torchgeometry's code).


def angle_axis_to_rotation_matrix(angle_axis: Tensor) -> Tensor:
    """torchgeometry 0.1.2 conversions.angle_axis_to_rotation_matrix: (n,3) -> (n,4,4)."""
    aa = angle_axis
    theta2 = (aa * aa).sum(dim=1, keepdim=True)
    theta = torch.sqrt(theta2)
    w = aa / (theta + 1e-6)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = torch.cos(theta), torch.sin(theta)
    normal = torch.cat([c + wx * wx * (1 - c), wx * wy * (1 - c) - wz * s, wy * s + wx * wz * (1 - c),
                        wz * s + wx * wy * (1 - c), c + wy * wy * (1 - c), -wx * s + wy * wz * (1 - c),
                        -wy * s + wx * wz * (1 - c), wx * s + wy * wz * (1 - c), c + wz * wz * (1 - c)],
                       dim=1).view(-1, 3, 3)
    rx, ry, rz = aa[:, 0:1], aa[:, 1:2], aa[:, 2:3]
    one = torch.ones_like(rx)
    taylor = torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)
    mask = (theta2 > 1e-6).view(-1, 1, 1).to(aa.dtype)
    out = torch.eye(4, dtype=aa.dtype).view(1, 4, 4).repeat(aa.shape[0], 1, 1)
    out[:, :3, :3] = mask * normal + (1 - mask) * taylor
    return out


def _member_lists(seg_label: Tensor, bb_labels: Optional[Tensor], K: int):
    """data_utils.py:1018-1025, :1038-1061: members[b][k] = ascending point indices with label k and bb == 0."""
    B, N = seg_label.shape
    out = []
    for b in range(B):
        row = []
        for k in range(K):
            m = seg_label[b] == k
            if bb_labels is not None:
                m = m & (bb_labels[b] == 0)
            row.append(m.nonzero().reshape(-1))
        out.append(row)
    return out


def sketch_implicit_projection(P, X, seg_label, bb_labels, axes, centers, S=1024, all_points=False,
                               zero_tol=1.0e-6):
    """data_utils.py:1014-1146 (variants 2/3: :1149-1417) -> P_proj (K,B,S,2), X_proj (K,B,S,2), scales (K,B),
    found (B,K).  Consumes the global CPU generator in the reference's order (:1064)."""
    B, K, _ = axes.shape
    N = P.shape[1]
    members = None if all_points else _member_lists(seg_label, bb_labels, K)
    P_proj = torch.zeros(K, B, S, 2)
    X_proj = torch.zeros(K, B, S, 2)
    found = torch.zeros(B, K)
    scales = torch.ones(K, B)
    z = torch.tensor([0.0, 0.0, 1.0])
    for i in range(K):
        cnt = [N if all_points else int(members[b][i].numel()) for b in range(B)]
        if sum(cnt) <= 1:
            continue
        pts = torch.zeros(B, S, 3)
        nrm = torch.zeros(B, S, 3)
        for j in range(B):
            if cnt[j] <= 1:
                continue
            if all_points:
                sel = torch.arange(N)
            else:
                sel = members[j][i][torch.randint(0, cnt[j], (S,))]
            pts[j], nrm[j] = P[j, sel], X[j, sel]
            found[j, i] = 1.0
        R = torch.eye(3).repeat(B, 1, 1)
        ang = torch.acos((axes[:, i] * z).sum(-1))
        for a in range(B):
            if ang[a] > zero_tol:
                rot_axis = torch.linalg.cross(axes[a, i], z)
                R[a] = angle_axis_to_rotation_matrix((rot_axis * ang[a]).unsqueeze(0))[0, :3, :3]
        pp = torch.bmm(pts, R)[:, :, :2]
        xp = torch.bmm(nrm, R)[:, :, :2]
        cp = torch.bmm(centers[:, i].unsqueeze(1), R)[:, :, :2]
        pp = pp - cp
        scales[i] = pp.pow(2).sum(-1).sqrt().max(dim=-1)[0]
        P_proj[i], X_proj[i] = pp, xp
    scales = torch.where(found.T == 1, scales, torch.ones_like(scales))
    return P_proj, X_proj, scales, found


def get_extrusion_extents(P, seg_label, bb_labels, axes, centers, S=1024):
    """data_utils.py:1650-1730 -> extents (K,B,2), found (B,K)."""
    B, K, _ = axes.shape
    members = _member_lists(seg_label, bb_labels, K)
    found = torch.zeros(B, K)
    extents = torch.zeros(K, B, 2)
    for i in range(K):
        cnt = [int(members[b][i].numel()) for b in range(B)]
        if sum(cnt) <= 1:
            continue
        pts = torch.zeros(B, S, 3)
        for j in range(B):
            if cnt[j] <= 1:
                continue
            pts[j] = P[j, members[j][i][torch.randint(0, cnt[j], (S,))]]
            found[j, i] = 1.0
        d = ((pts - centers[:, i].unsqueeze(1)) * axes[:, i].unsqueeze(1)).sum(-1)
        extents[i, :, 0], extents[i, :, 1] = d.min(dim=-1)[0], d.max(dim=-1)[0]
    return extents, found


def segment_centroids(EA_W: Tensor, pcs: Tensor):
    """eval.py:409-436: mean of the points with EA_W == 1 per (cloud, segment) when there are >= 2 of them."""
    B, N, K = EA_W.shape
    cen = torch.zeros(B, K, 3)
    found = torch.zeros(B, K)
    for j in range(K):
        for b in range(B):
            sel = (EA_W[b, :, j] == 1).nonzero().reshape(-1)
            if sel.numel() <= 1:
                continue
            cen[b, j] = pcs[b, sel].mean(dim=0)
            found[b, j] = 1.0
    return cen, found
