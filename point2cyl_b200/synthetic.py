"""Synthetic clouds for parity tests and benchmarks.

The reference's Fusion360 H5 dataset is a download (README.md:32) and is not available, so the
benchmarks use clouds of the same shape and labelling:

  s_cyl(B, N, K, seed)    extruded-cylinder scenes: per cloud 1..K instances, each a cylinder with a
                          random axis/centre/radius/half-extent, 70 % barrel points (normal
                          perpendicular to the axis, bb=0) and 30 % cap points (normal = +-axis, bb=1),
                          shuffled, centred and scaled into the unit sphere; labels are gap-free
                          0..n_inst-1 as losses.py:34 assumes.
  s_uniform(B, N, seed)   uniform points in [-1,1]^3 (sparse balls: exercises ball-query padding).

Everything is drawn from a seeded CPU torch.Generator so the oracle (CPU) and the CUDA path see
identical inputs.
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def _unit(v: torch.Tensor) -> torch.Tensor:
    return v / v.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def s_uniform(B: int, N: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, 3, generator=g) * 2 - 1


def s_cyl(B: int, N: int, K: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    pcs = torch.empty(B, N, 3)
    normals = torch.empty(B, N, 3)
    inst = torch.empty(B, N, dtype=torch.long)
    bb = torch.empty(B, N, dtype=torch.long)
    axes = torch.zeros(B, K, 3)
    centers = torch.zeros(B, K, 3)
    for b in range(B):
        n_inst = int(torch.randint(1, K + 1, (1,), generator=g))
        # split N as evenly as possible over the instances
        counts = [N // n_inst + (1 if i < N % n_inst else 0) for i in range(n_inst)]
        P, Nn, I, Bb = [], [], [], []
        ax_b = _unit(torch.randn(n_inst, 3, generator=g))
        c_b = torch.rand(n_inst, 3, generator=g) - 0.5
        rad = 0.1 + 0.3 * torch.rand(n_inst, generator=g)
        half = 0.1 + 0.3 * torch.rand(n_inst, generator=g)
        for i, cnt in enumerate(counts):
            a = ax_b[i]
            # orthonormal frame (u, v, a)
            helper = torch.tensor([1.0, 0.0, 0.0]) if abs(float(a[0])) < 0.9 else torch.tensor([0.0, 1.0, 0.0])
            u = _unit(torch.linalg.cross(a, helper))
            v = torch.linalg.cross(a, u)
            n_barrel = int(round(0.7 * cnt))
            n_cap = cnt - n_barrel
            th = torch.rand(n_barrel, generator=g) * (2 * math.pi)
            h = (torch.rand(n_barrel, generator=g) * 2 - 1) * half[i]
            radial = torch.cos(th)[:, None] * u + torch.sin(th)[:, None] * v
            P.append(c_b[i] + rad[i] * radial + h[:, None] * a)
            Nn.append(radial)
            Bb.append(torch.zeros(n_barrel, dtype=torch.long))
            th = torch.rand(n_cap, generator=g) * (2 * math.pi)
            rr = rad[i] * torch.sqrt(torch.rand(n_cap, generator=g))
            side = (torch.randint(0, 2, (n_cap,), generator=g) * 2 - 1).float()
            P.append(c_b[i] + rr[:, None] * (torch.cos(th)[:, None] * u + torch.sin(th)[:, None] * v)
                     + (side * half[i])[:, None] * a)
            Nn.append(side[:, None] * a.expand(n_cap, 3))
            Bb.append(torch.ones(n_cap, dtype=torch.long))
            I.append(torch.full((cnt,), i, dtype=torch.long))
        P, Nn, I, Bb = torch.cat(P), torch.cat(Nn), torch.cat(I), torch.cat(Bb)
        perm = torch.randperm(N, generator=g)
        P, Nn, I, Bb = P[perm], Nn[perm], I[perm], Bb[perm]
        # centre and scale into the unit sphere (what the reference's preprocessing does to a shape)
        mid = P.mean(dim=0, keepdim=True)
        P = P - mid
        scale = P.norm(dim=1).max().clamp_min(1e-12)
        P = P / scale
        pcs[b], normals[b], inst[b], bb[b] = P, Nn, I, Bb
        axes[b, :n_inst] = ax_b
        centers[b, :n_inst] = (c_b - mid) / scale
    return dict(pcs=pcs.float(), normals=normals.float(), inst=inst, bb=bb,
                axes=axes.float(), centers=centers.float())
