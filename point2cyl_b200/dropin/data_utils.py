"""Drop-in for the hot-path functions of the reference's data_utils.py: the closed-form extrusion fits.

estimate_extrusion_axis never builds the reference's (B,N,N) diag_embed matrices: BtB - CtC is the 3x3 scatter
sum_n (w_barrel^2 - w_base^2) x x^T, accumulated by p2c_segfit_stats_w, and the eigenvector of its smallest
eigenvalue comes from p2c_eig3x3_smallest (Jacobi, float64).  Differentiable: the backward kernels are
p2c_segfit_backward_w and p2c_eig3x3_backward (point2cyl_b200.autograd).
"""
import numpy as np
import torch

from point2cyl_b200 import autograd as ag
from point2cyl_b200 import ops
from point2cyl_b200.dropin.global_variables import *  # noqa: F401,F403


def add_noise(batch_xyz, batch_normal, sigma=0.01):
    """data_utils.py:84-96: host-side jitter of every point along its normal, before the device copy.  Same numpy
    RNG call and shape as the reference (one N(0, sigma) draw per point, (B, N)), same float64 promotion of the
    result (the training script casts to float32 when it moves the batch to the device)."""
    batch_xyz, batch_normal = torch.as_tensor(batch_xyz), torch.as_tensor(batch_normal)
    B, N, _ = batch_xyz.shape
    noise = torch.from_numpy(np.random.normal(0.0, sigma, (B, N)))
    return batch_xyz + noise[:, :, None] * batch_normal


def _sym3(m6):
    """(..., 6) [xx xy xz yy yz zz] -> (..., 3, 3)."""
    xx, xy, xz, yy, yz, zz = m6.unbind(-1)
    return torch.stack([torch.stack([xx, xy, xz], -1), torch.stack([xy, yy, yz], -1),
                        torch.stack([xz, yz, zz], -1)], -2)


def estimate_extrusion_axis(X, W_barrel, W_base, gt_bb_labels, gt_extrusion_instances, normalize=False):
    """data_utils.py:99-177 -> E_AX (B,K,3): unit eigenvector of the smallest eigenvalue of BtB - CtC per
    segment.  The eigenvector sign is arbitrary (the loss uses |dot|); here the largest component is positive."""
    B, N, K = W_barrel.shape
    st = ag.segfit_stats_w(W_barrel, W_base, X, False, None, None,
                            gt_extrusion_instances if normalize else None, gt_bb_labels if normalize else None)
    L = ops.seg_layout(K)
    Mb = st[:, L["Mbar"]:L["Mbar"] + 6 * K].reshape(B, K, 6)
    Mc = st[:, L["Mbase"]:L["Mbase"] + 6 * K].reshape(B, K, 6)
    if normalize:
        nb = torch.sqrt(st[:, L["cbar"]:L["cbar"] + K]) + 1.0
        nc = torch.sqrt(st[:, L["cbase"]:L["cbase"] + K]) + 1.0
        Mb = Mb / (nb * nb)[:, :, None]
        Mc = Mc / (nc * nc)[:, :, None]
    vec, _ = ag.eig3x3_smallest(_sym3(Mb - Mc))
    return vec


def estimate_extrusion_centers(W, pcs):
    """data_utils.py:253-266 -> (B,K,3): mean_n W[b,n,k] * p[b,n]  (a plain mean over N, not / sum W)."""
    B, N, K = W.shape
    st = ag.segfit_stats_w(W, None, None, False, pcs, None, None, None)
    L = ops.seg_layout(K)
    return st[:, L["C"]:L["C"] + 3 * K].reshape(B, K, 3) / N


# ---- projection / scale / extent closed forms -------------------------------------------------


def _draw_member_samples(counts, S):
    """The reference's random stream, data_utils.py:1064: for every segment i (outer) and cloud j (inner) that pass
    the two `<= 1 member` checks (:1042, :1054), one `torch.randint(0, n_members, (S,))` on the CPU generator.
    One device->host copy of the (B,K) member counts replaces the reference's K*B `nonzero()` syncs."""
    cnt = counts.cpu()
    B, K = cnt.shape
    rnd = torch.zeros(K, B, S, dtype=torch.long)
    tot = cnt.sum(dim=0)
    for i in range(K):
        if int(tot[i]) <= 1:
            continue
        for j in range(B):
            n = int(cnt[j, i])
            if n <= 1:
                continue
            rnd[i, j] = torch.randint(0, n, (S,))
    return rnd.to(counts.device, non_blocking=True)


def _project(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers, S, all_points):
    B, K, _ = extrusion_axes.shape
    N = P.shape[1]
    if all_points:
        # variant 3 (:1294): every point is a member of every segment and is used in order, so N must equal S
        if N != S:
            raise RuntimeError(f"sketch_implicit_projection3 needs num_points_to_sample == N ({S} vs {N})")
        counts = torch.full((B, K), N, dtype=torch.int32, device=P.device)
        lists = rnd = None
    else:
        counts, lists = ops.segment_lists(seg_label, bb_labels, 0, K)
        rnd = _draw_member_samples(counts, S)
    if torch.is_tensor(X) and X.requires_grad:
        # predicted normals (train_Point2Cyl.py:549): the projected normals carry a gradient back to X
        return ag.SketchProject.apply(X, P, lists, counts, rnd, extrusion_axes, extrusion_centers, S, g_zero_tol)  # noqa: F405
    return ops.sketch_project(P, X, lists, counts, rnd, extrusion_axes, extrusion_centers, S, g_zero_tol)  # noqa: F405


def sketch_implicit_projection(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers,
                               num_points_to_sample=1024):
    """data_utils.py:1014-1146 -> P_projected (K,B,S,2), X_projected (K,B,S,2), scales (K,B)."""
    Pp, Xp, sc, _ = _project(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers, num_points_to_sample, False)
    return Pp, Xp, sc


def sketch_implicit_projection2(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers,
                                num_points_to_sample=1024):
    """data_utils.py:1149-1281: as above plus found_centers_mask (B,K)."""
    return _project(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers, num_points_to_sample, False)


def sketch_implicit_projection3(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers,
                                num_points_to_sample=8192):
    """data_utils.py:1284-1417: all points of the cloud, unsampled, for every segment."""
    return _project(P, X, seg_label, bb_labels, extrusion_axes, extrusion_centers, num_points_to_sample, True)


def get_extrusion_extents(P, seg_label, bb_labels, extrusion_axes, extrusion_centers, num_points_to_sample=1024):
    """data_utils.py:1650-1730 -> extents (K,B,2) = min/max of (p - c).a over the sampled barrel points, found (B,K)."""
    B, K, _ = extrusion_axes.shape
    counts, lists = ops.segment_lists(seg_label, bb_labels, 0, K)
    rnd = _draw_member_samples(counts, num_points_to_sample)
    return ops.extrusion_extents(P, lists, counts, rnd, extrusion_axes, extrusion_centers, num_points_to_sample)


def estimate_segment_centroids(EA_W, pcs):
    """Function form of the inline centroid loop of eval.py:409-436: per (cloud, segment) the mean of the points with
    EA_W == 1 when there are at least two of them -> predicted_centroids (B,K,3), found_centers_mask (B,K)."""
    B, N, K = EA_W.shape
    hard = (EA_W == 1).float()
    st = ops.segfit_stats_w(hard, None, None, False, pcs, None, None, None)
    L = ops.seg_layout(K)
    cnt = st[:, L["colsum"]:L["colsum"] + K]
    found = cnt > 1.5
    cen = st[:, L["C"]:L["C"] + 3 * K].reshape(B, K, 3) / cnt.clamp_min(1.0)[:, :, None]
    return torch.where(found[:, :, None], cen, torch.zeros_like(cen)), found.float()
