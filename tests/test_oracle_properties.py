"""Definition-level properties of the oracle (independent of the goldens): each point operator is checked against a
brute-force statement of what the reference computes, on small random inputs.  The GPU parity tests compare the kernels
with this oracle; these tests make sure the oracle itself means what SURVEY.md section 8a says."""
import itertools

import numpy as np
import pytest
import torch

from oracle import p2c_oracle as orc


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fps_is_greedy_farthest_point(seed):
    """pointnet_util.py:63-84: start at `start`, then repeatedly the FIRST point of maximal distance to the chosen set."""
    g = torch.Generator().manual_seed(seed)
    B, N, S = 3, 200, 40
    xyz = torch.rand(B, N, 3, generator=g)
    start = torch.randint(0, N, (B,), generator=g)
    idx = orc.farthest_point_sample(xyz, S, start)
    assert idx.dtype == torch.long and idx.shape == (B, S)
    for b in range(B):
        assert int(idx[b, 0]) == int(start[b])
        assert len(set(idx[b].tolist())) == S                      # distinct points: never re-picks a chosen one
        dmin = torch.full((N,), 1e10)
        for j in range(S - 1):
            c = xyz[b, idx[b, j]]
            dmin = torch.minimum(dmin, ((xyz[b] - c) ** 2).sum(-1))
            assert int(idx[b, j + 1]) == int(torch.argmax(dmin))    # first maximum


@pytest.mark.parametrize("seed,radius,nsample", [(0, 0.25, 16), (1, 0.15, 8), (2, 0.6, 32)])
def test_ball_query_definition(seed, radius, nsample):
    """pointnet_util.py:87-107: ascending indices of the points within the radius, first `nsample`, padded with the
    first hit (the centre is itself a cloud point, so there is always one)."""
    g = torch.Generator().manual_seed(seed)
    B, N, S = 2, 300, 20
    xyz = torch.rand(B, N, 3, generator=g)
    centres = xyz[:, :S].clone()
    out = orc.query_ball_point(radius, nsample, xyz, centres)
    assert out.shape == (B, S, nsample) and out.dtype == torch.long
    d = orc.square_distance(centres, xyz)
    for b in range(B):
        for s in range(S):
            hits = torch.nonzero(d[b, s] <= radius ** 2).reshape(-1).tolist()
            want = hits[:nsample] + [hits[0]] * max(0, nsample - len(hits))
            assert out[b, s].tolist() == want


def test_three_nn_weights_and_neighbours():
    """pointnet_util.py:301-308: the three nearest sources, weights 1/(d+1e-8) normalised to one."""
    g = torch.Generator().manual_seed(3)
    B, N, S, D = 2, 50, 12, 5
    xyz1, xyz2 = torch.rand(B, N, 3, generator=g), torch.rand(B, S, 3, generator=g)
    feats = torch.randn(B, S, D, generator=g)
    out = orc.three_nn_interpolate(xyz1, xyz2, feats)
    out = out[0] if isinstance(out, tuple) else out
    d = ((xyz1[:, :, None, :] - xyz2[:, None, :, :]) ** 2).sum(-1)
    dd, ii = torch.sort(d, dim=-1)
    w = 1.0 / (dd[:, :, :3] + 1e-8)
    w = w / w.sum(-1, keepdim=True)
    ref = (torch.gather(feats[:, None].expand(B, N, S, D), 2, ii[:, :, :3, None].expand(B, N, 3, D)) * w[..., None]).sum(2)
    assert torch.allclose(out.reshape(ref.shape) if out.shape != ref.shape else out, ref, atol=1e-5)
    const = orc.three_nn_interpolate(xyz1, xyz2, torch.ones(B, S, 1))
    const = const[0] if isinstance(const, tuple) else const
    assert torch.allclose(const, torch.ones_like(const), atol=1e-6)    # partition of unity


@pytest.mark.parametrize("K", [2, 3, 5])
def test_hungarian_is_the_optimum_over_all_assignments(K):
    """losses.py:33-47: the matched columns maximise the summed relaxed IoU over the existing gt instances."""
    g = torch.Generator().manual_seed(K)
    B, N = 4, 64
    W = torch.softmax(torch.randn(B, N, K, generator=g) * 2, dim=-1)
    n_inst = [1 + (b % K) for b in range(B)]
    I_gt = torch.stack([torch.randint(0, n, (N,), generator=g) for n in n_inst])
    for b, n in enumerate(n_inst):
        I_gt[b, :n] = torch.arange(n)                                # every label present
    match, mask = orc.hungarian_matching(W, I_gt)
    for b, n in enumerate(n_inst):
        assert mask[b].tolist() == [True] * n + [False] * (K - n) and match[b, n:].tolist() == [0] * (K - n)
        onehot = torch.eye(n)[I_gt[b]]
        inter = onehot.t() @ W[b]
        score = inter / (onehot.sum(0)[:, None] + W[b].sum(0)[None, :] - inter).clamp(min=1e-10)
        best = max(sum(float(score[i, p[i]]) for i in range(n)) for p in itertools.permutations(range(K), n))
        got = sum(float(score[i, match[b, i]]) for i in range(n))
        assert abs(got - best) <= 1e-6


def test_axis_fit_recovers_a_cylinder_axis():
    """data_utils.py:99-177: barrel normals are perpendicular to the axis and base normals parallel to it, so the
    eigenvector of the smallest eigenvalue of sum w_bar^2 x x^T - sum w_base^2 x x^T is the axis (up to sign)."""
    g = torch.Generator().manual_seed(5)
    B, N, K = 2, 400, 2
    axes = torch.nn.functional.normalize(torch.randn(B, K, 3, generator=g), dim=-1)
    inst = torch.randint(0, K, (B, N), generator=g)
    bb = (torch.rand(B, N, generator=g) < 0.3).long()
    a = torch.gather(axes, 1, inst[:, :, None].expand(B, N, 3))
    r = torch.randn(B, N, 3, generator=g)
    perp = torch.nn.functional.normalize(r - (r * a).sum(-1, keepdim=True) * a, dim=-1)
    X = torch.where(bb[:, :, None] == 0, perp, a)
    onehot = torch.nn.functional.one_hot(inst, K).float()
    W_bar, W_base = onehot * (bb == 0)[:, :, None], onehot * (bb == 1)[:, :, None]
    E = orc.estimate_extrusion_axis(X, W_bar, W_base)
    E = E[0] if isinstance(E, tuple) else E
    assert float(((E * axes).sum(-1).abs() - 1).abs().max()) <= 1e-4
    c = orc.estimate_extrusion_centers(onehot, X)
    assert torch.allclose(c, torch.einsum("bnk,bnc->bkc", onehot, X) / N, atol=1e-6)   # plain mean over N (:253-266)
