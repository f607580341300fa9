"""ctypes front end of oracle/p2c_oracle_c.c (TEST INFRASTRUCTURE ONLY; see the header of the C file).

`build()` compiles the C restatement with gcc into oracle/libp2c_oracle.so (git-ignored, travels to the GPU box with
the snapshot like the product library); `__graft_entry__.build()` calls it.  Only tests/ import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "p2c_oracle_c.c")
LIB = os.path.join(HERE, "libp2c_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    tmp = f"{LIB}.{os.getpid()}.tmp"
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", tmp, SRC,
           "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed building the C oracle:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        _lib.p2c_oracle_fps.argtypes = [vp, vp, i32, i32, i32, vp]
        _lib.p2c_oracle_ball_query.argtypes = [vp, vp, i32, i32, i32, f32, i32, vp]
        _lib.p2c_oracle_three_nn.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
        for fn in (_lib.p2c_oracle_fps, _lib.p2c_oracle_ball_query, _lib.p2c_oracle_three_nn):
            fn.restype = None
    return _lib


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


def farthest_point_sample(xyz: torch.Tensor, npoint: int, start: torch.Tensor) -> torch.Tensor:
    """models/pointnet_util.py:63-84 -> (B, npoint) int64."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    start = start.detach().to("cpu", torch.long).contiguous()
    idx = torch.empty(B, npoint, dtype=torch.long)
    _load().p2c_oracle_fps(xyz.data_ptr(), start.data_ptr(), B, N, npoint, idx.data_ptr())
    return idx


def query_ball_point(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """models/pointnet_util.py:87-107 -> (B, S, nsample) int64; the threshold is float32(radius ** 2) like the
    reference's comparison against a Python double."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty(B, S, nsample, dtype=torch.long)
    r2 = float(np.float32(float(radius) ** 2))
    _load().p2c_oracle_ball_query(xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, r2, nsample, idx.data_ptr())
    return idx


def three_nn(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """models/pointnet_util.py:301-306 -> (idx (B,N,3) int64, weight (B,N,3), dist (B,N,3))."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S < 3:
        raise ValueError("three_nn needs at least three source points")
    idx = torch.empty(B, N, 3, dtype=torch.long)
    w = torch.empty(B, N, 3, dtype=torch.float32)
    d = torch.empty(B, N, 3, dtype=torch.float32)
    _load().p2c_oracle_three_nn(xyz1.data_ptr(), xyz2.data_ptr(), B, N, S, idx.data_ptr(), w.data_ptr(), d.data_ptr())
    return idx, w, d
