"""Drop-in for the LIVE functions of the reference's losses.py (same names, signatures, return shapes).

All reductions over the N points run in the libp2c.so statistics kernel (p2c_segfit_stats_w), the assignment in
p2c_hungarian; what is left in torch are (B,K)-sized elementwise formulas.  Results carry an autograd graph whose backward runs
the p2c_segfit_backward_w kernel (point2cyl_b200.autograd).  Dead reference code (sketch / chamfer / axis-regularisation losses,
losses.py:165-312) is intentionally absent.
"""
import math

import torch

from point2cyl_b200 import autograd as ag
from point2cyl_b200 import ops
from point2cyl_b200.dropin.global_variables import *  # noqa: F401,F403  (the reference re-exports these)

TORCH_PI = math.pi


def _stats(W, I_gt=None, X=None, X_gt=None):
    K = W.shape[2]
    st = ag.segfit_stats_w(W, None, X, False, None, X_gt, I_gt, None)
    return st, ops.seg_layout(K), K


def hungarian_matching(W_pred, I_gt, with_mask=False):
    """losses.py:22-52 -> matching_indices (B,K) int64 [, mask (B,K) bool].  No gradient (like the reference)."""
    st, L, K = _stats(W_pred.detach(), I_gt)
    cost, n_gt = ops.segfit_cost(st.detach(), K)
    match = ops.hungarian(cost, n_gt)
    if not with_mask:
        return match
    mask = torch.arange(K, device=W_pred.device)[None, :] < n_gt[:, None]
    return match, mask


def hard_W_encoding(W, to_null_mask=False, W_null_threshold=0.005):
    """losses.py:55-68: arg-max one-hot, columns with sum W < threshold*N zeroed when asked (p2c_hard_w_encoding)."""
    B, N, K = W.shape
    colsum = None
    if to_null_mask:
        st, L, _ = _stats(W)
        colsum = st[:, L["colsum"]:L["colsum"] + K]
    return ops.hard_w_encoding(W, colsum, float(N) * W_null_threshold)


def sequence_mask(lengths, maxlen=None):
    """losses.py:70-76."""
    if maxlen is None:
        maxlen = int(lengths.max())
    return torch.arange(0, maxlen, 1, device=lengths.device) < lengths.unsqueeze(-1)


def get_mask_gt(I_gt, n_max_instances):
    """losses.py:78-81."""
    return sequence_mask(torch.max(I_gt, dim=1)[0] + 1, maxlen=n_max_instances)


def reduce_mean_masked_instance(loss, mask_gt):
    """losses.py:83-88."""
    loss = torch.where(mask_gt, loss, torch.zeros_like(loss))
    denom = mask_gt.float().sum(dim=1)
    return torch.where(denom > 0, loss.sum(dim=1) / denom, torch.zeros_like(denom))


def compute_miou_loss(W, I_gt, matching_indices, div_eps=1e-10):
    """losses.py:90-103 -> (1 - mIoU (B,K), 1 - dot/N (B,K), W_reordered (B,N,K))."""
    B, N, K = W.shape
    L_ = matching_indices.shape[1]
    st, L, _ = _stats(W, I_gt)
    D = st[:, :K * K].reshape(B, K, K)                               # D[g, k] = sum_n [I_gt = g] W[n, k]
    dot = torch.gather(D[:, :L_, :], 2, matching_indices[:, :, None]).squeeze(2)
    cnt = st[:, L["cnt"]:L["cnt"] + L_]
    colsum = torch.gather(st[:, L["colsum"]:L["colsum"] + K], 1, matching_indices)
    miou = dot / (cnt + colsum - dot + div_eps)
    W_reordered = torch.gather(W, 2, matching_indices.unsqueeze(1).expand(B, N, L_))
    return 1.0 - miou, 1 - dot / N, W_reordered


def compute_segmentation_iou(W, I_gt, matching_indices, mask):
    """losses.py:106-109."""
    mIoU = 1 - compute_miou_loss(W, I_gt, matching_indices)[0]
    return torch.sum(mask * mIoU, dim=1) / torch.sum(mask, dim=1)


def acos_safe(x):
    """losses.py:123-124."""
    return torch.acos(torch.clamp(x, min=-1.0 + 1e-6, max=1.0 - 1e-6))


def compute_normal_loss(normal, normal_gt, angle_diff, collapse=True):
    """losses.py:127-143 (unoriented normals)."""
    if not angle_diff and collapse and normal.dim() == 3 and normal.shape[1] >= 1024:
        # mean_n (1 - |<x, x_gt>|) straight from the statistics kernel
        B, N, _ = normal.shape
        dummy = normal.new_zeros(B, N, 1)
        st = ag.segfit_stats_w(dummy, None, normal, False, None, normal_gt, None, None)
        return st[:, ops.seg_layout(1)["normal"]] / N
    dot_abs = torch.abs(torch.sum(normal * normal_gt, dim=2))
    val = acos_safe(dot_abs) if angle_diff else 1.0 - dot_abs
    return torch.mean(val, dim=1) if collapse else val


def compute_normal_difference(X, X_gt, in_radians=True, collapse=True):
    """losses.py:146-159 (p2c_normal_angle)."""
    scale = 1.0 if in_radians else 180.0 / TORCH_PI
    return ops.normal_angle(X, X_gt, scale, collapse)


def compute_all_losses(P, W, I_gt, X, X_gt, normal_loss_multiplier, miou_loss_multiplier,
                       return_match_indices=False, collapse=True):
    """losses.py:317-351."""
    B, N, K = W.shape
    mask_gt = get_mask_gt(I_gt, K)
    st = ag.segfit_stats_w(W, None, X, False, None, X_gt, I_gt, None)   # one pass: D, counts, sums, normal loss
    L = ops.seg_layout(K)
    if normal_loss_multiplier > 0:
        normal_loss = st[:, L["normal"]] / N
    else:
        normal_loss = torch.zeros([B, K], device=P.device)
    matching_indices = mask = None
    if miou_loss_multiplier > 0:
        cost, n_gt = ops.segfit_cost(st.detach(), K)
        matching_indices = ops.hungarian(cost, n_gt)
        mask = torch.arange(K, device=W.device)[None, :] < n_gt[:, None]
        D = st[:, :K * K].reshape(B, K, K)
        dot = torch.gather(D, 2, matching_indices[:, :, None]).squeeze(2)
        colsum = torch.gather(st[:, L["colsum"]:L["colsum"] + K], 1, matching_indices)
        miou_loss = 1.0 - dot / (st[:, L["cnt"]:L["cnt"] + K] + colsum - dot + 1e-10)
        avg_miou_loss = reduce_mean_masked_instance(miou_loss, mask_gt)
    else:
        avg_miou_loss = torch.zeros([B, K], device=P.device)
    if collapse:
        total_miou_loss = torch.mean(avg_miou_loss)
        total_normal_loss = torch.mean(normal_loss)
    else:
        total_miou_loss, total_normal_loss = avg_miou_loss, normal_loss
    total_loss = miou_loss_multiplier * total_miou_loss + normal_loss_multiplier * total_normal_loss
    if return_match_indices:
        return total_loss, total_normal_loss, total_miou_loss, matching_indices, mask
    return total_loss, total_normal_loss, total_miou_loss
