"""Pins the C restatement of the integer-valued point operators (oracle/p2c_oracle_c.c) to the reference goldens and to
the torch oracle, then uses it where the torch oracle cannot go: BASELINE.json's full and stress sizes on the CPU."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import c_oracle as corc
from oracle import p2c_oracle as orc
from point2cyl_b200 import synthetic

POINTOPS = [("pointops_cyl_n1024.npz", "cyl"), ("pointops_uniform_n2048.npz", "uniform")]


def inputs(golden_dir, name, kind):
    g = np.load(os.path.join(golden_dir, name))
    B, N, npoint, nsample, seed = (int(v) for v in g["meta"])
    xyz = synthetic.s_cyl(B, N, 4, seed)["pcs"] if kind == "cyl" else synthetic.s_uniform(B, N, seed)
    return g, xyz, npoint, float(g["radius"]), nsample


@pytest.mark.parametrize("name,kind", POINTOPS)
def test_c_oracle_reproduces_the_reference_goldens(golden_dir, name, kind):
    g, xyz, npoint, radius, nsample = inputs(golden_dir, name, kind)
    fps = corc.farthest_point_sample(xyz, npoint, torch.from_numpy(g["start"]))
    assert np.array_equal(fps.numpy(), g["fps_idx"].astype(np.int64))
    new_xyz = orc.gather_points(xyz, fps)
    grp = corc.query_ball_point(radius, nsample, xyz, new_xyz)
    assert np.array_equal(grp.numpy(), g["group_idx"].astype(np.int64))
    idx, w, d = corc.three_nn(xyz, new_xyz)
    assert np.array_equal(idx.numpy(), g["nn_idx"].astype(np.int64))
    assert np.array_equal(w.numpy(), g["nn_w"])                       # the interpolation weights too, bit for bit


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_c_oracle_equals_torch_oracle(seed):
    """Random clouds, a lattice (many exactly equal distances -> tie rules) and duplicated points."""
    g = torch.Generator().manual_seed(seed)
    B, N, S = 3, 700, 96
    xyz = torch.rand(B, N, 3, generator=g) * 2 - 1
    if seed == 2:
        xyz = (xyz * 4).round() / 4                                   # lattice: ties everywhere
    if seed == 3:
        xyz[:, 100:200] = xyz[:, :100]                                # exact duplicates
    start = torch.randint(0, N, (B,), generator=g)
    fps = corc.farthest_point_sample(xyz, S, start)
    assert torch.equal(fps, orc.farthest_point_sample(xyz, S, start))
    new_xyz = orc.gather_points(xyz, fps)
    for radius, nsample in ((0.2, 16), (0.45, 64), (0.01, 8)):
        assert torch.equal(corc.query_ball_point(radius, nsample, xyz, new_xyz),
                           orc.query_ball_point(radius, nsample, xyz, new_xyz))
    idx, w, d = corc.three_nn(xyz, new_xyz)
    ref_d, ref_order = orc.square_distance(xyz, new_xyz).sort(dim=-1, stable=True)
    assert torch.equal(d, ref_d[:, :, :3])                            # the three smallest distances, bit for bit
    if seed < 2:                                                      # generic positions: the indices too
        assert torch.equal(idx, ref_order[:, :, :3])


def test_ball_without_hits_yields_N():
    xyz = torch.rand(1, 50, 3)
    far = torch.full((1, 2, 3), 10.0)
    out = corc.query_ball_point(0.1, 4, xyz, far)
    assert torch.equal(out, torch.full((1, 2, 4), 50, dtype=torch.long))
    assert torch.equal(out, orc.query_ball_point(0.1, 4, xyz, far))


def test_full_size_properties_config2_and_stress():
    """The sizes the torch oracle cannot reach on a CPU in test time: B=4 clouds of config 2 (N=8192) and of the stress
    configuration (N=32768, S-uniform).  Checked through size-independent properties: distinct FPS picks that start at
    `start`, every ball member within the radius and ascending up to the padding, padding = first hit, and the first
    cloud against the torch oracle."""
    for N, kind, S, radius, nsample in ((8192, "cyl", 512, 0.2, 64), (32768, "uniform", 512, 0.2, 64)):
        B = 4
        xyz = synthetic.s_cyl(B, N, 8, 1234)["pcs"] if kind == "cyl" else synthetic.s_uniform(B, N, 7)
        start = torch.arange(B) * 17
        t0 = time.perf_counter()
        fps = corc.farthest_point_sample(xyz, S, start)
        new_xyz = orc.gather_points(xyz, fps)
        grp = corc.query_ball_point(radius, nsample, xyz, new_xyz)
        assert time.perf_counter() - t0 < 60
        assert torch.equal(fps[:, 0], start)
        assert all(len(set(fps[b].tolist())) == S for b in range(B))
        pts = orc.gather_points(xyz, grp)                              # (B,S,ns,3)
        d = ((pts - new_xyz[:, :, None, :]) ** 2).sum(-1)
        assert float(d.max()) <= radius ** 2 * (1 + 1e-4) + 1e-6
        first = grp[:, :, :1]
        inc = (grp[:, :, 1:] > grp[:, :, :-1]) | (grp[:, :, 1:] == first)   # ascending, then padded with the first
        assert bool(inc.all())
        assert torch.equal(fps[:1], orc.farthest_point_sample(xyz[:1], S, start[:1]))
        assert torch.equal(grp[:1, :64], orc.query_ball_point(radius, nsample, xyz[:1], new_xyz[:1, :64]))
