"""Drop-in for the hot-path functions of the reference's data_utils.py: the closed-form extrusion fits.

estimate_extrusion_axis never builds the reference's (B,N,N) diag_embed matrices: BtB - CtC is the 3x3 scatter
sum_n (w_barrel^2 - w_base^2) x x^T, accumulated by p2c_segfit_stats_w, and the eigenvector of its smallest
eigenvalue comes from p2c_eig3x3_smallest (Jacobi, float64).  Forward only this round.
"""
import numpy as np
import torch

from point2cyl_b200 import ops
from point2cyl_b200.dropin.global_variables import *  # noqa: F401,F403


def add_noise(pcs, normals, sigma=0.01):
    """data_utils.py:84-96: host-side jitter along the normals (numpy RNG, before the device copy)."""
    pcs_np = pcs.numpy() if torch.is_tensor(pcs) else np.asarray(pcs)
    nrm_np = normals.numpy() if torch.is_tensor(normals) else np.asarray(normals)
    noise = np.random.normal(0.0, sigma, size=pcs_np.shape[:-1] + (1,)).astype(pcs_np.dtype)
    out = pcs_np + noise * nrm_np
    return torch.from_numpy(out) if torch.is_tensor(pcs) else out


def _sym3(m6):
    """(..., 6) [xx xy xz yy yz zz] -> (..., 3, 3)."""
    xx, xy, xz, yy, yz, zz = m6.unbind(-1)
    return torch.stack([torch.stack([xx, xy, xz], -1), torch.stack([xy, yy, yz], -1),
                        torch.stack([xz, yz, zz], -1)], -2)


def estimate_extrusion_axis(X, W_barrel, W_base, gt_bb_labels, gt_extrusion_instances, normalize=False):
    """data_utils.py:99-177 -> E_AX (B,K,3): unit eigenvector of the smallest eigenvalue of BtB - CtC per
    segment.  The eigenvector sign is arbitrary (the loss uses |dot|); here the largest component is positive."""
    B, N, K = W_barrel.shape
    st = ops.segfit_stats_w(W_barrel, W_base, X, False, None, None,
                            gt_extrusion_instances if normalize else None, gt_bb_labels if normalize else None)
    L = ops.seg_layout(K)
    Mb = st[:, L["Mbar"]:L["Mbar"] + 6 * K].reshape(B, K, 6)
    Mc = st[:, L["Mbase"]:L["Mbase"] + 6 * K].reshape(B, K, 6)
    if normalize:
        nb = torch.sqrt(st[:, L["cbar"]:L["cbar"] + K]) + 1.0
        nc = torch.sqrt(st[:, L["cbase"]:L["cbase"] + K]) + 1.0
        Mb = Mb / (nb * nb)[:, :, None]
        Mc = Mc / (nc * nc)[:, :, None]
    vec, _ = ops.eig3x3_smallest(_sym3(Mb - Mc))
    return vec


def estimate_extrusion_centers(W, pcs):
    """data_utils.py:253-266 -> (B,K,3): mean_n W[b,n,k] * p[b,n]  (a plain mean over N, not / sum W)."""
    B, N, K = W.shape
    st = ops.segfit_stats_w(W, None, None, False, pcs, None, None, None)
    L = ops.seg_layout(K)
    return st[:, L["C"]:L["C"] + 3 * K].reshape(B, K, 3) / N
