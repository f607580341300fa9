"""Thin tensor-level wrappers over the C-ABI (one function per kernel entry point).

Layout: clouds (B,N,3) float32 contiguous; feature matrices are 2-D (rows, C) float32 with
stride(1) == 1 and an arbitrary row stride (so a column slice of a concat buffer is a valid
operand); indices int64.  Everything runs on torch's current CUDA stream.  No autograd here —
point2cyl_b200.autograd wraps these for training.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, need_cuda, ptr, stream_ptr

Tensor = torch.Tensor


def _rows(t: Tensor) -> Tensor:
    if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.float32:
        raise _lib.P2CError(f"expected a 2-D float32 matrix with unit column stride, got "
                            f"{tuple(t.shape)} strides {t.stride()} {t.dtype}")
    return t


def _cloud(t: Tensor) -> Tensor:
    if t.dim() != 3 or t.shape[2] != 3:
        raise _lib.P2CError(f"expected (B,N,3), got {tuple(t.shape)}")
    return t.contiguous().float()


def _out(out: Optional[Tensor], shape, dtype, device) -> Tensor:
    """A caller-provided result buffer (static across CUDA-graph replays) or a fresh one."""
    if out is None:
        return torch.empty(*shape, dtype=dtype, device=device)
    if tuple(out.shape) != tuple(shape) or out.dtype != dtype or not out.is_contiguous() or out.device != device:
        raise _lib.P2CError(f"out= buffer must be contiguous {tuple(shape)} {dtype}, got {tuple(out.shape)} {out.dtype}")
    return out


def set_sm_budget(sms: int) -> int:
    """SMs the persistent tensor-core kernels of later launches may occupy (0 = all); returns the previous value."""
    return int(_lib.load().p2c_set_sm_budget(int(sms)))


def fps(xyz: Tensor, npoint: int, start: Tensor, out: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """(idx (B,npoint) int64, new_xyz (B,npoint,3)).  start: (B,) int64 first centroid per cloud."""
    need_cuda(xyz)
    xyz = _cloud(xyz)
    B, N, _ = xyz.shape
    start = start.to(device=xyz.device, dtype=torch.long).contiguous()
    idx = _out(None if out is None else out[0], (B, npoint), torch.long, xyz.device)
    new_xyz = _out(None if out is None else out[1], (B, npoint, 3), torch.float32, xyz.device)
    call("p2c_fps", ptr(xyz), ptr(start), B, N, npoint, ptr(idx), ptr(new_xyz), stream_ptr())
    return idx, new_xyz


def radius_sq_f32(radius: float) -> float:
    """float32(radius**2): the threshold the reference's `sqrdists > radius ** 2` compares against."""
    return float(np.float32(float(radius) ** 2))


def ball_query(radius: float, nsample: int, xyz: Tensor, new_xyz: Tensor, out: Optional[Tensor] = None) -> Tensor:
    need_cuda(xyz, new_xyz)
    xyz, new_xyz = _cloud(xyz), _cloud(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = _out(out, (B, S, nsample), torch.long, xyz.device)
    call("p2c_ball_query", ptr(xyz), ptr(new_xyz), B, N, S, radius_sq_f32(radius), nsample,
                                     ptr(out), stream_ptr())
    return out


def pad4(c: int) -> int:
    return (c + 3) // 4 * 4


def group(xyz: Tensor, feats: Optional[Tensor], new_xyz: Optional[Tensor], idx: Optional[Tensor],
          ldo: Optional[int] = None) -> Tensor:
    """Grouped rows (B*S*nsample, ldo): [xyz[idx]-centre, feats[idx]], zero padded.
    feats: (B*N, D) rows.  idx None = group-all (one group per cloud, centre 0)."""
    need_cuda(xyz, feats, new_xyz, idx)
    xyz = _cloud(xyz)
    B, N, _ = xyz.shape
    D = 0 if feats is None else _rows(feats).shape[1]
    if idx is None:
        S, ns = 1, N
    else:
        idx = idx.contiguous()
        S, ns = idx.shape[1], idx.shape[2]
        new_xyz = _cloud(new_xyz)
    ldo = ldo or pad4(3 + D)
    out = torch.empty(B * S * ns, ldo, dtype=torch.float32, device=xyz.device)
    call("p2c_group", ptr(xyz), ptr(feats), 0 if feats is None else feats.stride(0),
                                ptr(new_xyz) if idx is not None else None, ptr(idx), B, N, S, ns, D,
                                ptr(out), ldo, stream_ptr())
    return out


def sa_first_layer(xyz: Tensor, new_xyz: Tensor, idx: Tensor, Qf: Optional[Tensor], W: Tensor,
                   bias: Optional[Tensor], stats: Optional[Tensor]) -> Tensor:
    """Fused grouping gather + first SA conv.  W (C, 3+D[,1,1]); Qf = feats @ W[:, 3:].T rows or None (D = 0)."""
    xyz, new_xyz = _cloud(xyz), _cloud(new_xyz)
    B, N, _ = xyz.shape
    idx = idx.contiguous()
    S, ns = idx.shape[1], idx.shape[2]
    W2 = W.reshape(W.shape[0], -1)
    if not W2.is_contiguous():
        W2 = W2.contiguous()
    C_ = W2.shape[0]
    Y = torch.empty(B * S * ns, C_, dtype=torch.float32, device=xyz.device)
    call("p2c_sa_first_layer", ptr(xyz), ptr(new_xyz), ptr(idx), ptr(Qf), 0 if Qf is None else Qf.stride(0), ptr(W2),
         W2.stride(0), ptr(bias), B, N, S, ns, C_, ptr(Y), Y.stride(0), ptr(stats), stream_ptr())
    return Y


def linear_small(X: Tensor, W: Tensor, bias: Optional[Tensor]) -> Tensor:
    """Y = X W^T + bias for a handful of rows (p2c_linear_small: M <= 256, K <= 3072, fp32).  W (N, K) may be a column
    slice of a wider matrix."""
    X = _rows(X)
    W2 = W.reshape(W.shape[0], -1) if W.dim() != 2 else W
    if W2.stride(1) != 1:
        W2 = W2.contiguous()
    M, K = X.shape
    N = W2.shape[0]
    if W2.shape[1] != K:
        raise _lib.P2CError(f"linear_small: X {tuple(X.shape)} W {tuple(W2.shape)}")
    Y = torch.empty(M, N, dtype=torch.float32, device=X.device)
    call("p2c_linear_small", ptr(X), X.stride(0), ptr(W2), W2.stride(0), ptr(bias), ptr(Y), N, M, N, K, stream_ptr())
    return Y


def linear_group_bias(X: Tensor, w_split: Tensor, bias_rows: Tensor, group: int, N: int, K: int,
                      in_scale: Optional[Tensor] = None, in_shift: Optional[Tensor] = None,
                      in_bn: Optional["PendingBN"] = None, stats: Optional[Tensor] = None) -> Tensor:
    """Y[m] = f(X[m]) W^T + bias_rows[m // group] on the streamed-weight tensor-core kernel (p2c_linear_group_bias).
    w_split (2, N, pad4(K)) from split_tf32[_multi]; bias_rows (M / group, N) contiguous."""
    X = _rows(X)
    M = X.shape[0]
    bias_rows = _rows(bias_rows)
    if bias_rows.shape != (-(-M // group), N) or not bias_rows.is_contiguous():
        raise _lib.P2CError(f"linear_group_bias: bias_rows {tuple(bias_rows.shape)} for M={M}, group={group}, N={N}")
    Y = torch.empty(M, N, dtype=torch.float32, device=X.device)
    ps, ph, d = _fold_args(in_bn, in_scale, in_shift)
    call("p2c_linear_group_bias", ptr(X), X.stride(0), ptr(w_split), w_split.shape[-1], ptr(bias_rows), group, ps, ph, d,
         ptr(Y), N, M, N, K, ptr(stats), stream_ptr())
    return Y


def group_moments(xyz: Tensor, new_xyz: Tensor, idx: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """Per-CTA partial sums of the nine first / second moments of the centred neighbour coordinates
    xyz[b, idx[b,s,j]] - new_xyz[b,s] over all rows (coordinates only: geometry stage); input of `sa_xyz_stats`."""
    xyz, new_xyz = _cloud(xyz), _cloud(new_xyz)
    B, N, _ = xyz.shape
    idx = idx.contiguous()
    n = int(_lib.load().p2c_group_moments_size())
    # (the last word is the kernel's launch counter: zero before the first launch, left zero by every launch)
    out = torch.zeros(n, dtype=torch.float64, device=xyz.device) if out is None else _out(out, (n,), torch.float64, xyz.device)
    call("p2c_group_moments", ptr(xyz), ptr(new_xyz), ptr(idx), B, N, idx.shape[1], idx.shape[2], ptr(out), stream_ptr())
    return out


def sa_xyz_stats(partials: Tensor, rows: int, W: Tensor, bias: Optional[Tensor], stats: Tensor) -> None:
    """BatchNorm sum / sum-of-squares of the xyz-only first SA conv in closed form (p2c_sa_xyz_stats)."""
    W2 = W.reshape(W.shape[0], -1)
    if not W2.is_contiguous():
        W2 = W2.contiguous()
    call("p2c_sa_xyz_stats", ptr(partials), rows, ptr(W2), W2.stride(0), ptr(bias), W2.shape[0], ptr(stats), stream_ptr())


def sa_xyz_linear(xyz: Tensor, new_xyz: Tensor, idx: Tensor, W0: Tensor, b0: Optional[Tensor], W1: Tensor,
                  b1: Optional[Tensor], scale0: Optional[Tensor] = None, shift0: Optional[Tensor] = None,
                  bn0: Optional["PendingBN"] = None, stats: Optional[Tensor] = None, pool_group: int = 0,
                  want_y: bool = True, moments: Optional[Tensor] = None):
    """Second layer of a feature-less SA level with the first (xyz-only) layer recomputed in its operand transform
    (p2c_sa_xyz_linear).  Returns Y (rows, N1) or, with pool_group, (Y or None, Ymax, Ymin); None when the tensor-core
    kernel does not take the shape (the caller then materialises the first layer)."""
    xyz, new_xyz = _cloud(xyz), _cloud(new_xyz)
    B, N, _ = xyz.shape
    idx = idx.contiguous()
    S, ns = idx.shape[1], idx.shape[2]
    W0_ = W0.reshape(W0.shape[0], -1)
    W1_ = W1.reshape(W1.shape[0], -1)
    if not W0_.is_contiguous():
        W0_ = W0_.contiguous()
    if not W1_.is_contiguous():
        W1_ = W1_.contiguous()
    C0, N1 = W0_.shape[0], W1_.shape[0]
    if W1_.shape[1] != C0 or W0_.shape[1] != 3:
        raise _lib.P2CError(f"sa_xyz_linear: W0 {tuple(W0_.shape)} / W1 {tuple(W1_.shape)}")
    rows = B * S * ns
    Y = torch.empty(rows, N1, dtype=torch.float32, device=xyz.device) if want_y else None
    Ymax = Ymin = None
    if pool_group:
        Ymax = torch.empty(rows // pool_group, N1, dtype=torch.float32, device=xyz.device)
        Ymin = torch.empty_like(Ymax)
    pending = bn0 is not None and bn0.pending
    ps, ph, d = _fold_args(bn0, scale0, shift0)
    try:
        call("p2c_sa_xyz_linear", ptr(xyz), ptr(new_xyz), ptr(idx), B, N, S, ns, ptr(W0_), W0_.stride(0), ptr(b0), C0,
             ps, ph, d, ptr(moments) if d is not None else None, ptr(W1_), ptr(b1), N1, ptr(Y), 0 if Y is None else Y.stride(0), ptr(stats), pool_group,
             ptr(Ymax), ptr(Ymin), stream_ptr())
    except _lib.P2CError as e:
        if "P2C_EUNSUPPORTED" not in str(e):
            raise
        if pending:
            bn0.pending = True
        return None
    if pool_group:
        return Y, Ymax, Ymin
    return Y


def sa_stack_fused(xyz: Tensor, new_xyz: Tensor, idx: Tensor, convs, bns: "list[PendingBN]",
                   out: Optional[Tensor] = None) -> Optional[Tensor]:
    """A whole feature-less SA level in ONE kernel (p2c_sa_stack_fused): gather + three conv/BN/ReLU layers + max over
    the neighbours, for BatchNorms whose statistics are known up front (eval mode).  convs: the level's three 1x1
    convs, bns: their PendingBN descriptors.  -> pooled post-BN/ReLU rows (B*S, C2), or None when the kernel does not
    take the shape (the caller then runs the per-layer path)."""
    need_cuda(xyz, new_xyz, idx)
    xyz, new_xyz = _cloud(xyz), _cloud(new_xyz)
    B, N, _ = xyz.shape
    S, ns = idx.shape[1], idx.shape[2]
    W = [c.weight.reshape(c.weight.shape[0], -1) for c in convs]
    C0, C1, C2 = (w.shape[0] for w in W)
    if (len(convs) != 3 or W[0].shape[1] != 3 or C0 != 64 or C1 != 64 or C2 > 128 or ns not in (32, 64, 128) or
            B * S * ns >= 2 ** 31):
        return None
    if out is None:
        out = torch.empty(B * S, C2, dtype=torch.float32, device=xyz.device)
    W = [w if w.is_contiguous() else w.contiguous() for w in W]
    call("p2c_sa_stack_fused", ptr(xyz.contiguous()), ptr(new_xyz.contiguous()), ptr(idx.contiguous()), B, N, S, ns,
         ptr(W[0]), W[0].stride(0), ptr(convs[0].bias), bns[0].fold(), ptr(W[1]), ptr(convs[1].bias), bns[1].fold(),
         ptr(W[2]), ptr(convs[2].bias), bns[2].fold(), C0, C1, C2, ptr(out), out.stride(0), stream_ptr())
    return out


def linear(X: Tensor, W: Tensor, bias: Optional[Tensor], K: Optional[int] = None,
           in_scale: Optional[Tensor] = None, in_shift: Optional[Tensor] = None,
           in_mask: Optional[Tensor] = None, stats: Optional[Tensor] = None, pool_group: int = 0,
           out: Optional[Tensor] = None, want_y: bool = True, precision: int = _lib.PREC_FP32,
           w_split: Optional[Tensor] = None, in_bn: Optional["PendingBN"] = None):
    """One MLP layer.  X (M, >=K) rows, W (N, K) (any trailing singleton dims), returns Y (M,N) or,
    with pool_group, (Y or None, Ymax, Ymin).  stats: float64 (2N,) accumulator (pre-zeroed)."""
    need_cuda(X, W)
    X = _rows(X)
    W2 = W.reshape(W.shape[0], -1)
    if not W2.is_contiguous():
        W2 = W2.contiguous()
    N = W2.shape[0]
    K = K or W2.shape[1]
    if W2.shape[1] != K or X.shape[1] < K:
        raise _lib.P2CError(f"linear: K mismatch X{tuple(X.shape)} W{tuple(W2.shape)} K={K}")
    M = X.shape[0]
    Y = None
    if want_y:
        Y = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=X.device)
        _rows(Y)
    Ymax = Ymin = None
    if pool_group:
        Ymax = torch.empty(M // pool_group, N, dtype=torch.float32, device=X.device)
        Ymin = torch.empty_like(Ymax)
    if in_mask is not None:
        _rows(in_mask)
    ps, ph, d = _fold_args(in_bn, in_scale, in_shift)
    call("p2c_linear",
        ptr(X), X.stride(0), ptr(W2), ptr(bias), ps, ph, ptr(in_mask),
        0 if in_mask is None else in_mask.stride(0), ptr(Y), 0 if Y is None else Y.stride(0), M, N, K,
        ptr(stats), pool_group, ptr(Ymax), ptr(Ymin), precision, ptr(w_split),
        0 if w_split is None else w_split.shape[-1], d, stream_ptr())
    if pool_group:
        return Y, Ymax, Ymin
    return Y


def _seed(seed: Optional[Tensor]) -> Optional[Tensor]:
    if seed is not None and (seed.dtype != torch.long or seed.numel() != 2 or not seed.is_cuda
                             or not seed.is_contiguous()):
        raise _lib.P2CError("dropout seed: a contiguous CUDA int64 tensor of two words expected")
    return seed


def head_masked(H: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], mask_cf: Optional[Tensor],
                W: Tensor, bias: Optional[Tensor], B: int, N: int, seed: Optional[Tensor] = None,
                bn: Optional["PendingBN"] = None, precision: int = _lib.PREC_3XTF32) -> Tensor:
    """Output heads: (B*N, C) raw fc1 rows -> (B*N, Nout); mask_cf is the (B, C, N) dropout mask, or - with
    mask_cf None - `seed` (two int64 words on the device) makes the kernel draw the p=0.5 mask itself (Philox).
    precision 3xtf32 (default): the tcgen05 layer kernel when a BatchNorm is folded and no explicit mask is given;
    fp32 / explicit mask: the SIMT kernel."""
    H = _rows(H)
    C_ = H.shape[1]
    W2 = W.reshape(W.shape[0], -1).contiguous()
    Nout = W2.shape[0]
    if mask_cf is not None and (tuple(mask_cf.shape) != (B, C_, N) or not mask_cf.is_contiguous()):
        raise _lib.P2CError(f"head_masked: mask must be contiguous (B,C,N)=({B},{C_},{N}), got {tuple(mask_cf.shape)}")
    Y = torch.empty(B * N, Nout, dtype=torch.float32, device=H.device)
    ps, ph, d = _fold_args(bn, scale, shift)
    call("p2c_head_masked", ptr(H), H.stride(0), ps, ph, ptr(mask_cf), ptr(_seed(seed)), ptr(W2), ptr(bias), ptr(Y),
         Nout, B, N, C_, Nout, d, precision, stream_ptr())
    return Y


def split_tf32(W: Tensor) -> Tensor:
    """(2, N, pad4(K)) hi/lo copy of a weight matrix for the streamed-weight tensor-core kernel."""
    W2 = W.reshape(W.shape[0], -1)
    if not W2.is_contiguous():
        W2 = W2.contiguous()
    N, K = W2.shape
    out = torch.empty(2, N, pad4(K), dtype=torch.float32, device=W.device)
    call("p2c_split_tf32", ptr(W2), N, K, ptr(out), out.shape[2], stream_ptr())
    return out


def split_tf32_multi(weights, outs=None, transposed=None):
    """hi/lo split of several weight matrices in ONE launch.  weights: list of (N, K[,1[,1]]) tensors; outs: matching
    list of (2, N, pad4(K)) buffers to refresh (allocated when None).  transposed[i]: weights[i] is stored as (K, N) and
    the split of its transpose is produced (no transposed copy is made).  Returns the list of split buffers."""
    Ws = []
    for W in weights:
        W2 = W.reshape(W.shape[0], -1) if W.dim() != 2 else W
        Ws.append(W2 if W2.stride(1) == 1 else W2.contiguous())       # a column block of a wider matrix is fine
    tr = [bool(t) for t in transposed] if transposed is not None else [False] * len(Ws)
    shape = [(w.shape[1], w.shape[0]) if t else (w.shape[0], w.shape[1]) for w, t in zip(Ws, tr)]   # (N, K) of the result
    if outs is None:
        outs = [torch.empty(2, n, pad4(k), dtype=torch.float32, device=w.device) for w, (n, k) in zip(Ws, shape)]
    res = list(outs)
    for c0 in range(0, len(Ws), 16):
        ws, os_, sh, ts = Ws[c0:c0 + 16], outs[c0:c0 + 16], shape[c0:c0 + 16], tr[c0:c0 + 16]
        n = len(ws)
        pW = (C.c_void_p * n)(*[ptr(w) for w in ws])
        pO = (C.c_void_p * n)(*[ptr(o) for o in os_])
        aN = (C.c_int * n)(*[x[0] for x in sh])
        aK = (C.c_int * n)(*[x[1] for x in sh])
        aL = (C.c_int64 * n)(*[o.shape[2] for o in os_])
        aT = (C.c_int * n)(*[1 if t else 0 for t in ts])
        aS = (C.c_int64 * n)(*[w.stride(0) for w in ws])
        call("p2c_split_tf32_multi", pW, aN, aK, pO, aL, aT, aS, n, stream_ptr())
    return res


def linear_act(X: Tensor, w_split: Tensor, bias: Optional[Tensor], N: int, K: int, op: int, beta: float = 100.0,
               oscale: float = 1.0, out: Optional[Tensor] = None, S: Optional[Tensor] = None,
               mul: Optional[Tensor] = None) -> Tensor:
    """One implicit-network layer on the tensor cores (p2c_linear_act): op 1 = softplus (+ sigmoid into S),
    op 2 = multiply by `mul`, op 0 = plain.  w_split: (2, N, pad4(K)) from split_tf32[_multi]."""
    need_cuda(X, w_split)
    X = _rows(X)
    M = X.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=X.device)
    _rows(out)
    if S is not None:
        _rows(S)
    if mul is not None:
        _rows(mul)
    call("p2c_linear_act", ptr(X), X.stride(0), ptr(w_split), w_split.shape[-1], ptr(bias), M, N, K, op, float(beta),
         float(oscale), ptr(out), out.stride(0), ptr(S), 0 if S is None else S.stride(0), ptr(mul),
         0 if mul is None else mul.stride(0), stream_ptr())
    return out


def linear_act_bwd(X: Tensor, w_split: Tensor, N: int, K: int, op: int, mul: Tensor, v: Tensor, out: Tensor,
                   beta: float = 100.0, oscale: float = 1.0, Z: Optional[Tensor] = None) -> Tensor:
    """One layer-GEMM of the implicit network's second-order backward (p2c_linear_act_bwd): op 3 writes
    out = acc * mul * oscale and Z = beta * acc * v * (1 - mul); op 4 writes out = acc * mul * oscale + v."""
    need_cuda(X, w_split, mul, v, out)
    X = _rows(X)
    for t in (mul, v, out) + ((Z,) if Z is not None else ()):
        _rows(t)
    M = X.shape[0]
    call("p2c_linear_act_bwd", ptr(X), X.stride(0), ptr(w_split), w_split.shape[-1], M, N, K, op, float(beta),
         float(oscale), ptr(out), out.stride(0), ptr(mul), mul.stride(0), ptr(v), v.stride(0), ptr(Z),
         0 if Z is None else Z.stride(0), stream_ptr())
    return out


def cast_bf16(W: Tensor) -> Tensor:
    """(N, pad8(K)) bf16 copy of a weight matrix for the bf16 tensor-core path (P2C_PREC_BF16)."""
    W2 = W.reshape(W.shape[0], -1)
    if not W2.is_contiguous():
        W2 = W2.contiguous()
    N, K = W2.shape
    out = torch.empty(N, (K + 7) // 8 * 8, dtype=torch.bfloat16, device=W.device)
    call("p2c_cast_bf16", ptr(W2), N, K, ptr(out), out.shape[1], stream_ptr())
    return out


def weight_operand(X: Tensor, W: Tensor, N: int, K: int, has_mask: bool, pool_group: int,
                   precision: int) -> Optional[Tensor]:
    """The pre-processed weight copy p2c_linear wants as `w_split` for this layer: hi/lo tf32 split (streamed-weight
    3xTF32 kernel), bf16 copy (bf16 kernel) or None."""
    path = _lib.load().p2c_linear_path(X.stride(0), X.shape[0], N, K, int(has_mask), pool_group, precision, 1)
    if path == 2:
        return split_tf32(W)
    if path == 3:
        return cast_bf16(W)
    return None


def needs_split(X: Tensor, N: int, K: int, has_mask: bool, pool_group: int, precision: int) -> bool:
    """True when p2c_linear would use the streamed-weight kernel for this layer if given a split copy."""
    return _lib.load().p2c_linear_path(X.stride(0), X.shape[0], N, K, int(has_mask), pool_group, precision, 1) == 2


class PendingBN:
    """A BatchNorm whose finalisation is deferred to the kernel that consumes it (include/point2cyl.h: p2c_bn_fold).
    scale / shift (/ mean / invstd) are its output tensors: valid - in stream order - after the first consumer
    (`fold()` hands the descriptor to that kernel, once) or after `resolve()` (the stand-alone p2c_bn_finalize)."""

    def __init__(self, stats: Optional[Tensor], count: int, gamma: Tensor, beta: Tensor, eps: float, momentum: float,
                 training: bool, running_mean: Optional[Tensor], running_var: Optional[Tensor], save: bool = False,
                 out: Optional[Tensor] = None):
        C_ = gamma.shape[0]
        buf = out if out is not None else torch.empty(4 * C_, dtype=torch.float32, device=gamma.device)
        self.scale, self.shift = buf[:C_], buf[C_:2 * C_]
        self.mean = buf[2 * C_:3 * C_] if save else None
        self.invstd = buf[3 * C_:4 * C_] if save else None
        self.training, self.count, self.C = training, int(count), C_
        self._keep = (stats, gamma, beta, running_mean, running_var, buf)
        self._desc = _lib.BnFold(
            stats=ptr(stats) if training else None, count=int(count), gamma=ptr(gamma), beta=ptr(beta), eps=float(eps),
            momentum=float(momentum), running_mean=ptr(running_mean), running_var=ptr(running_var),
            scale_out=ptr(self.scale), shift_out=ptr(self.shift), mean_out=ptr(self.mean),
            invstd_out=ptr(self.invstd), C=C_)
        self.pending = True

    def fold(self):
        """ctypes pointer to the descriptor for the ONE kernel that folds it (None once it has been consumed)."""
        if not self.pending:
            return None
        self.pending = False
        return C.byref(self._desc)

    def resolve(self) -> None:
        """Finalise with the stand-alone kernel if no consumer has folded it yet."""
        if not self.pending:
            return
        self.pending = False
        stats, gamma, beta, rm, rv, _ = self._keep
        call("p2c_bn_finalize", ptr(stats), self.count, ptr(gamma), ptr(beta), self._desc.eps, self._desc.momentum,
             1 if self.training else 0, ptr(rm), ptr(rv), ptr(self.scale), ptr(self.shift), ptr(self.mean),
             ptr(self.invstd), self.C, stream_ptr())


def _fold_args(bn: Optional["PendingBN"], scale: Optional[Tensor], shift: Optional[Tensor]):
    """(scale ptr, shift ptr, descriptor or None) for a consumer call: a still-pending BatchNorm goes in as the
    descriptor, an already finalised one (or plain arrays) as pointers."""
    d = bn.fold() if bn is not None else None
    if bn is not None and d is None:
        scale, shift = bn.scale, bn.shift
    if d is not None:
        return None, None, d
    return ptr(scale), ptr(shift), None


def bn_finalize(stats: Optional[Tensor], count: int, gamma: Tensor, beta: Tensor, eps: float,
                momentum: float, training: bool, running_mean: Optional[Tensor],
                running_var: Optional[Tensor], save: bool = False):
    """(scale, shift[, mean, invstd]) and in-place running-stat update when training."""
    C_ = gamma.shape[0]
    dev = gamma.device
    scale = torch.empty(C_, dtype=torch.float32, device=dev)
    shift = torch.empty(C_, dtype=torch.float32, device=dev)
    mean = torch.empty(C_, dtype=torch.float32, device=dev) if save else None
    invstd = torch.empty(C_, dtype=torch.float32, device=dev) if save else None
    call("p2c_bn_finalize", ptr(stats), count, ptr(gamma), ptr(beta), eps, momentum,
                                      1 if training else 0, ptr(running_mean), ptr(running_var),
                                      ptr(scale), ptr(shift), ptr(mean), ptr(invstd), C_, stream_ptr())
    if save:
        return scale, shift, mean, invstd
    return scale, shift


def bn_relu_apply(Y: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], out: Optional[Tensor] = None,
                  bn: Optional[PendingBN] = None) -> Tensor:
    Y = _rows(Y)
    M, C_ = Y.shape
    if out is None:
        out = torch.empty(M, C_, dtype=torch.float32, device=Y.device)
    _rows(out)
    ps, ph, d = _fold_args(bn, scale, shift)
    call("p2c_bn_relu_apply", ptr(Y), Y.stride(0), ps, ph, ptr(out), out.stride(0), M, C_, d, stream_ptr())
    return out


def pool_bn_relu(Ymax: Tensor, Ymin: Tensor, scale: Optional[Tensor], shift: Optional[Tensor],
                 out: Optional[Tensor] = None, bn: Optional[PendingBN] = None) -> Tensor:
    G, C_ = Ymax.shape
    if out is None:
        out = torch.empty(G, C_, dtype=torch.float32, device=Ymax.device)
    _rows(out)
    ps, ph, d = _fold_args(bn, scale, shift)
    call("p2c_pool_bn_relu", ptr(Ymax), ptr(Ymin), ps, ph, ptr(out), out.stride(0), G, C_, d, stream_ptr())
    return out


def three_nn_interp(xyz1: Tensor, xyz2: Tensor, feats2: Tensor, out: Optional[Tensor] = None,
                    want_idx: bool = False):
    """feats2 (B*S, D) rows -> out (B*N, D) rows (may be a column slice of a wider buffer)."""
    need_cuda(xyz1, xyz2, feats2)
    xyz1, xyz2 = _cloud(xyz1), _cloud(xyz2)
    feats2 = _rows(feats2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    D = feats2.shape[1]
    if out is None:
        out = torch.empty(B * N, D, dtype=torch.float32, device=xyz1.device)
    _rows(out)
    idx = w = None
    if want_idx and S > 1:
        idx = torch.empty(B, N, 3, dtype=torch.long, device=xyz1.device)
        w = torch.empty(B, N, 3, dtype=torch.float32, device=xyz1.device)
    call("p2c_three_nn_interp", ptr(xyz1), ptr(xyz2), ptr(feats2), feats2.stride(0), B, N, S, D,
                                          ptr(out), out.stride(0), ptr(idx), ptr(w), stream_ptr())
    if want_idx:
        return out, idx, w
    return out


def three_nn_search(xyz1: Tensor, xyz2: Tensor, out: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """The coordinate-only half of three_nn_interp: (idx (B,N,3) int64, w (B,N,3)) of the three nearest sources."""
    need_cuda(xyz1, xyz2)
    xyz1, xyz2 = _cloud(xyz1), _cloud(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = _out(None if out is None else out[0], (B, N, 3), torch.long, xyz1.device)
    w = _out(None if out is None else out[1], (B, N, 3), torch.float32, xyz1.device)
    call("p2c_three_nn_search", ptr(xyz1), ptr(xyz2), B, N, S, ptr(idx), ptr(w), stream_ptr())
    return idx, w


def three_nn_gather(feats2: Tensor, idx: Tensor, w: Tensor, S: int, out: Optional[Tensor] = None) -> Tensor:
    """The feature half: out (B*N, D) rows = sum_j w[b,n,j] * feats2[b*S + idx[b,n,j]] (same rounding as the fused
    kernel).  `out` may be a column slice of a wider buffer."""
    need_cuda(feats2, idx, w)
    feats2 = _rows(feats2)
    B, N, _ = idx.shape
    D = feats2.shape[1]
    if not (idx.is_contiguous() and w.is_contiguous() and idx.dtype == torch.long and w.dtype == torch.float32):
        raise _lib.P2CError("three_nn_gather: contiguous int64 idx / float32 w of shape (B,N,3) expected")
    if feats2.shape[0] != B * S:
        raise _lib.P2CError(f"three_nn_gather: feats2 has {feats2.shape[0]} rows, expected B*S = {B * S}")
    if out is None:
        out = torch.empty(B * N, D, dtype=torch.float32, device=feats2.device)
    _rows(out)
    call("p2c_three_nn_gather", ptr(feats2), feats2.stride(0), ptr(idx), ptr(w), B, N, S, D, ptr(out), out.stride(0),
         stream_ptr())
    return out


def segfit_stride(K: int) -> int:
    return _lib.load().p2c_segfit_stats_stride(K)


def segfit_stats(X_raw: Tensor, W_raw: Tensor, pcs: Tensor, gt_normals: Tensor, inst: Tensor,
                 bb: Tensor, K: int) -> Tensor:
    """Per-cloud sufficient statistics (B, stride(K)).  X_raw (B,N,>=3 strided rows), W_raw (B,N,2K)."""
    B, N = inst.shape
    Xr = _rows(X_raw.reshape(B * N, -1) if X_raw.dim() == 3 and X_raw.is_contiguous() else _as_rows(X_raw))
    Wr = _rows(W_raw.reshape(B * N, -1) if W_raw.dim() == 3 and W_raw.is_contiguous() else _as_rows(W_raw))
    pcs, gt_normals = _cloud(pcs), _cloud(gt_normals)
    inst, bb = inst.contiguous(), bb.contiguous()
    stride = segfit_stride(K)
    nchunks = (N + 255) // 256
    partial = torch.empty(B * nchunks * stride, dtype=torch.float32, device=pcs.device)
    stats = torch.empty(B, stride, dtype=torch.float32, device=pcs.device)
    call("p2c_segfit_stats", ptr(Xr), Xr.stride(0), ptr(Wr), Wr.stride(0), ptr(pcs),
                                       ptr(gt_normals), ptr(inst), ptr(bb), B, N, K, ptr(partial),
                                       partial.numel(), ptr(stats), stream_ptr())
    return stats


def _w_operand(t: Optional[Tensor], B: int, N: int):
    """(pointer tensor, row stride, element stride) of a (B,N,K) soft-assignment tensor, copying only if its
    batch dimension does not collapse into rows."""
    if t is None:
        return None, 0, 0
    if t.dtype != torch.float32 or t.stride(0) != N * t.stride(1):
        t = t.contiguous().float()
    return t, t.stride(1), t.stride(2)


def segfit_stats_w(Wb: Tensor, Wc: Optional[Tensor] = None, X: Optional[Tensor] = None, normalize_x: bool = False,
                   pcs: Optional[Tensor] = None, gt_normals: Optional[Tensor] = None,
                   inst: Optional[Tensor] = None, bb: Optional[Tensor] = None) -> Tensor:
    """Per-cloud statistics (layout of p2c_segfit_stats) from soft assignments (B,N,K) the caller holds."""
    need_cuda(Wb)
    B, N, K = Wb.shape
    wb, ldb, sb = _w_operand(Wb, B, N)
    wc, ldc, sc = _w_operand(Wc, B, N)
    Xr = None
    if X is not None:
        Xr = _rows(_as_rows(X.float()))
    pcs = None if pcs is None else _cloud(pcs)
    gt_normals = None if gt_normals is None else _cloud(gt_normals)
    inst = None if inst is None else inst.contiguous().long()
    bb = None if bb is None else bb.contiguous().long()
    stride = segfit_stride(K)
    nchunks = (N + 255) // 256
    partial = torch.empty(B * nchunks * stride, dtype=torch.float32, device=Wb.device)
    stats = torch.empty(B, stride, dtype=torch.float32, device=Wb.device)
    call("p2c_segfit_stats_w", ptr(Xr), 0 if Xr is None else Xr.stride(0), 1 if normalize_x else 0, ptr(wb), ldb, sb,
         ptr(wc), ldc, sc, ptr(pcs), ptr(gt_normals), ptr(inst), ptr(bb), B, N, K, ptr(partial), partial.numel(),
         ptr(stats), stream_ptr())
    return stats


def seg_layout(K: int):
    """Offsets into a stats row (see csrc/segfit.cu)."""
    return dict(D=0, cnt=K * K, cbar=K * K + K, cbase=K * K + 2 * K, colsum=K * K + 3 * K, C=K * K + 4 * K,
                Mbar=K * K + 7 * K, Mbase=K * K + 13 * K, normal=K * K + 19 * K, maxlab=K * K + 19 * K + 1)


def _as_rows(t: Tensor) -> Tensor:
    """(B,N,C) tensor whose (B,N) dims collapse to one row dim with unit column stride, else copy."""
    B, N, C_ = t.shape
    if t.stride(2) == 1 and t.stride(0) == N * t.stride(1):
        return t.as_strided((B * N, C_), (t.stride(1), 1), t.storage_offset())
    return t.contiguous().reshape(B * N, C_)


def segfit_cost(stats: Tensor, K: int) -> Tuple[Tensor, Tensor]:
    B = stats.shape[0]
    cost = torch.empty(B, K, K, dtype=torch.float32, device=stats.device)
    n_gt = torch.empty(B, dtype=torch.int32, device=stats.device)
    call("p2c_segfit_cost", ptr(stats), B, K, ptr(cost), ptr(n_gt), stream_ptr())
    return cost, n_gt


def hungarian(cost: Tensor, n_gt: Tensor) -> Tensor:
    """cost (B,K,K) scores, n_gt (B) int32 -> match (B,K) int64, on the device (no host sync)."""
    B, K, _ = cost.shape
    match = torch.empty(B, K, dtype=torch.long, device=cost.device)
    call("p2c_hungarian", ptr(cost.contiguous()), ptr(n_gt), B, K, ptr(match), stream_ptr())
    return match


def bb_loss_sums(W_raw: Tensor, bb: Tensor, match: Tensor, n_gt: Tensor, K: int) -> Tensor:
    B, N = bb.shape
    Wr = _rows(_as_rows(W_raw))
    nchunks = (N + 255) // 256
    partial = torch.empty(B * nchunks, dtype=torch.float32, device=bb.device)
    out = torch.empty(B, dtype=torch.float32, device=bb.device)
    call("p2c_bb_loss", ptr(Wr), Wr.stride(0), ptr(bb.contiguous()), ptr(match.contiguous()),
                                  ptr(n_gt), B, N, K, ptr(partial), partial.numel(), ptr(out), stream_ptr())
    return out


def loss_finalize(stats: Tensor, bb_sum: Optional[Tensor], match: Tensor, n_gt: Tensor, gt_axes: Tensor,
                  gt_centers: Tensor, N: int, K: int, norm_eig: bool, weights):
    B = stats.shape[0]
    dev = stats.device
    E_AX = torch.empty(B, K, 3, dtype=torch.float32, device=dev)
    centers = torch.empty(B, K, 3, dtype=torch.float32, device=dev)
    per_seg = torch.empty(B, K, 3, dtype=torch.float32, device=dev)
    per_cloud = torch.empty(B, 5, dtype=torch.float32, device=dev)
    losses = torch.empty(6, dtype=torch.float32, device=dev)
    w = (C.c_float * 5)(*[float(v) for v in weights])
    call("p2c_loss_finalize", ptr(stats), ptr(bb_sum), ptr(match.contiguous()), ptr(n_gt),
                                        ptr(gt_axes.contiguous().float()), ptr(gt_centers.contiguous().float()),
                                        B, N, K, 1 if norm_eig else 0, w, ptr(E_AX), ptr(centers),
                                        ptr(per_seg), ptr(per_cloud), ptr(losses), stream_ptr())
    return losses, E_AX, centers, per_seg, per_cloud


def eig3x3_smallest(M: Tensor) -> Tuple[Tensor, Tensor]:
    """M (..., 3, 3) symmetric (upper triangle read) -> (vec (...,3), eigenvalues ascending (...,3))."""
    need_cuda(M)
    flat = M.reshape(-1, 9).contiguous().float()
    n = flat.shape[0]
    vec = torch.empty(n, 3, dtype=torch.float32, device=M.device)
    ev = torch.empty(n, 3, dtype=torch.float32, device=M.device)
    call("p2c_eig3x3_smallest", ptr(flat), n, ptr(vec), ptr(ev), stream_ptr())
    return vec.reshape(*M.shape[:-2], 3), ev.reshape(*M.shape[:-2], 3)


def square_distance(src: Tensor, dst: Tensor) -> Tensor:
    need_cuda(src, dst)
    src, dst = _cloud(src), _cloud(dst)
    B, S, _ = src.shape
    N = dst.shape[1]
    out = torch.empty(B, S, N, dtype=torch.float32, device=src.device)
    call("p2c_square_distance", ptr(src), ptr(dst), B, S, N, ptr(out), stream_ptr())
    return out


def gather_rows(points: Tensor, idx: Tensor) -> Tensor:
    """points (B,N,C), idx (B, ...) int64 -> (B, ..., C)."""
    need_cuda(points, idx)
    B, N, C_ = points.shape
    pts = points if (points.stride(2) == 1 and points.stride(0) == N * points.stride(1)
                     and points.dtype == torch.float32) else points.contiguous().float()
    flat = idx.reshape(B, -1).contiguous()
    Mper = flat.shape[1]
    out = torch.empty(B, Mper, C_, dtype=torch.float32, device=points.device)
    call("p2c_gather_rows", ptr(pts), pts.stride(1), ptr(flat), B, N, Mper, C_, ptr(out), stream_ptr())
    return out.reshape(*idx.shape, C_)


# ---- second wave: projections / extents / eval helpers ----------------------------------


def segment_lists(seg_label: Tensor, bb: Optional[Tensor], bb_value: int, K: int, want_lists: bool = True):
    """counts (B,K) int32 [, lists (B,K,N) int32]: members of (b,k) are points with seg_label == k (and bb == bb_value),
    ascending point index."""
    need_cuda(seg_label, bb)
    B, N = seg_label.shape
    seg_label = seg_label.contiguous().long()
    bb = None if bb is None else bb.contiguous().long()
    counts = torch.empty(B, K, dtype=torch.int32, device=seg_label.device)
    lists = torch.empty(B, K, N, dtype=torch.int32, device=seg_label.device) if want_lists else None
    call("p2c_segment_lists", ptr(seg_label), ptr(bb), bb_value, B, N, K, ptr(counts), ptr(lists), stream_ptr())
    return counts, lists


def sketch_project(P: Tensor, X: Optional[Tensor], lists: Optional[Tensor], counts: Tensor, rand_idx: Optional[Tensor],
                   axes: Tensor, centers: Tensor, S: int, zero_tol: float, want_sel: bool = False):
    """-> P_proj (K,B,S,2), X_proj (K,B,S,2) | None, scales (K,B), found (B,K) [, sel (K,B,S) int32, R (K,B,9)]."""
    need_cuda(P, X, counts, axes, centers)
    P = _cloud(P)
    X = None if X is None else _cloud(X)
    B, N, _ = P.shape
    K = axes.shape[1]
    dev = P.device
    P_proj = torch.empty(K, B, S, 2, dtype=torch.float32, device=dev)
    X_proj = torch.empty(K, B, S, 2, dtype=torch.float32, device=dev) if X is not None else None
    scales = torch.empty(K, B, dtype=torch.float32, device=dev)
    found = torch.empty(B, K, dtype=torch.float32, device=dev)
    sel = torch.empty(K, B, S, dtype=torch.int32, device=dev) if want_sel else None
    R = torch.empty(K, B, 9, dtype=torch.float32, device=dev) if want_sel else None
    if rand_idx is not None:
        rand_idx = rand_idx.to(device=dev, dtype=torch.long).contiguous()
    call("p2c_sketch_project", ptr(P), ptr(X), B, N, K, S, ptr(lists), ptr(counts), ptr(rand_idx),
         ptr(axes.contiguous().float()), ptr(centers.contiguous().float()), float(zero_tol), ptr(P_proj), ptr(X_proj),
         ptr(scales), ptr(found), ptr(sel), ptr(R), stream_ptr())
    if want_sel:
        return P_proj, X_proj, scales, found, sel, R
    return P_proj, X_proj, scales, found


def sketch_project_bwd(dX_proj: Tensor, sel: Tensor, R: Tensor, B: int, N: int) -> Tensor:
    K, _, S = sel.shape
    dX = torch.empty(B, N, 3, dtype=torch.float32, device=sel.device)
    call("p2c_sketch_project_bwd", ptr(dX_proj.contiguous().float()), ptr(sel), ptr(R), B, N, K, S, ptr(dX), stream_ptr())
    return dX


def extrusion_extents(P: Tensor, lists: Optional[Tensor], counts: Tensor, rand_idx: Optional[Tensor], axes: Tensor,
                      centers: Tensor, S: int):
    """-> extents (K,B,2) [min, max of (p-c).a], found (B,K)."""
    need_cuda(P, counts, axes, centers)
    P = _cloud(P)
    B, N, _ = P.shape
    K = axes.shape[1]
    extents = torch.empty(K, B, 2, dtype=torch.float32, device=P.device)
    found = torch.empty(B, K, dtype=torch.float32, device=P.device)
    if rand_idx is not None:
        rand_idx = rand_idx.to(device=P.device, dtype=torch.long).contiguous()
    call("p2c_extrusion_extents", ptr(P), B, N, K, S, ptr(lists), ptr(counts), ptr(rand_idx),
         ptr(axes.contiguous().float()), ptr(centers.contiguous().float()), ptr(extents), ptr(found), stream_ptr())
    return extents, found


def hard_w_encoding(W: Tensor, colsum: Optional[Tensor], null_below: float, want_hard: bool = True,
                    want_label: bool = False):
    """W (B,N,K) (strided ok) -> hard (B,N,K) one-hot of the arg-max with nulled columns zeroed [, label (B,N)]."""
    need_cuda(W)
    B, N, K = W.shape
    w, ldw, sw = _w_operand(W, B, N)
    hard = torch.empty(B, N, K, dtype=torch.float32, device=W.device) if want_hard else None
    label = torch.empty(B, N, dtype=torch.long, device=W.device) if want_label else None
    if colsum is not None:
        colsum = colsum.contiguous().float()
    call("p2c_hard_w_encoding", ptr(w), ldw, sw, B, N, K, ptr(colsum), float(null_below), ptr(hard), ptr(label),
         stream_ptr())
    if want_hard and want_label:
        return hard, label
    return hard if want_hard else label


def normal_angle(X: Tensor, G: Tensor, scale: float = 1.0, collapse: bool = True) -> Tensor:
    """scale * acos_safe(|<x,g>|): per point (B,N) or its mean over N (B,)."""
    need_cuda(X, G)
    X, G = _cloud(X), _cloud(G)
    B, N, _ = X.shape
    if collapse:
        out = torch.empty(B, dtype=torch.float32, device=X.device)
        call("p2c_normal_angle", ptr(X), ptr(G), B, N, float(scale), None, ptr(out), stream_ptr())
        return out / N
    out = torch.empty(B, N, dtype=torch.float32, device=X.device)
    call("p2c_normal_angle", ptr(X), ptr(G), B, N, float(scale), ptr(out), None, stream_ptr())
    return out


# ---- backward of the loss block -------------------------------------------------------------------------------


def loss_backward_coef(stats: Tensor, match: Tensor, n_gt: Tensor, gt_axes: Tensor, gt_centers: Tensor,
                       eff_weights: Tensor, N: int, K: int, norm_eig: bool) -> Tensor:
    """d total / d stats (B, stride).  eff_weights: device (5,) float32 {seg, normal, bb, axis, centre}."""
    B = stats.shape[0]
    dstats = torch.empty_like(stats)
    call("p2c_loss_backward_coef", ptr(stats), ptr(match.contiguous()), ptr(n_gt), ptr(gt_axes.contiguous().float()),
         ptr(gt_centers.contiguous().float()), ptr(eff_weights), B, N, K, 1 if norm_eig else 0, ptr(dstats),
         stream_ptr())
    return dstats


def segfit_backward(X_raw: Tensor, W_raw: Tensor, pcs: Tensor, gt_normals: Tensor, inst: Tensor, bb: Tensor,
                    dstats: Tensor, match: Optional[Tensor], n_gt: Optional[Tensor], eff_weights: Optional[Tensor],
                    K: int, out: Optional[Tensor] = None) -> Tensor:
    """-> d(out) rows (B*N, 3 + 2K): gradient w.r.t. [X_raw | W_raw] in the head's output layout."""
    B, N = inst.shape
    Xr = _rows(_as_rows(X_raw))
    Wr = _rows(_as_rows(W_raw))
    if out is None:
        out = torch.empty(B * N, 3 + 2 * K, dtype=torch.float32, device=pcs.device)
    dX, dW = out[:, :3], out[:, 3:]
    call("p2c_segfit_backward", ptr(Xr), Xr.stride(0), ptr(Wr), Wr.stride(0), ptr(_cloud(pcs)), ptr(_cloud(gt_normals)),
         ptr(inst.contiguous()), ptr(bb.contiguous()), ptr(dstats), ptr(None if match is None else match.contiguous()),
         ptr(n_gt), ptr(eff_weights), B, N, K, ptr(dX), out.stride(0), ptr(dW), out.stride(0), stream_ptr())
    return out


def segfit_backward_w(dstats: Tensor, Wb: Tensor, Wc: Optional[Tensor], X: Optional[Tensor], normalize_x: bool,
                      pcs: Optional[Tensor], gt_normals: Optional[Tensor], inst: Optional[Tensor],
                      want_dx: bool, want_dwc: bool):
    """Backward of segfit_stats_w -> (dX (B,N,3) | None, dWb (B,N,K), dWc (B,N,K) | None)."""
    B, N, K = Wb.shape
    wb, ldb, sb = _w_operand(Wb, B, N)
    wc, ldc, sc = _w_operand(Wc, B, N)
    Xr = None if X is None else _rows(_as_rows(X.float()))
    dev = Wb.device
    dX = torch.empty(B, N, 3, dtype=torch.float32, device=dev) if (want_dx and Xr is not None) else None
    dWb = torch.empty(B, N, K, dtype=torch.float32, device=dev)
    dWc = torch.empty(B, N, K, dtype=torch.float32, device=dev) if (want_dwc and wc is not None) else None
    call("p2c_segfit_backward_w", ptr(Xr), 0 if Xr is None else Xr.stride(0), 1 if normalize_x else 0, ptr(wb), ldb, sb,
         ptr(wc), ldc, sc, ptr(None if pcs is None else _cloud(pcs)),
         ptr(None if gt_normals is None else _cloud(gt_normals)),
         ptr(None if inst is None else inst.contiguous().long()), ptr(dstats.contiguous()), B, N, K, ptr(dX), 3,
         ptr(dWb), ptr(dWc), stream_ptr())
    return dX, dWb, dWc


def eig3x3_backward(M: Tensor, gvec: Tensor) -> Tensor:
    flat = M.reshape(-1, 9).contiguous().float()
    g = gvec.reshape(-1, 3).contiguous().float()
    dM = torch.empty_like(flat)
    call("p2c_eig3x3_backward", ptr(flat), ptr(g), flat.shape[0], ptr(dM), stream_ptr())
    return dM.reshape(M.shape)


# ---- backward of the backbone (kernels in csrc/backward.cu) ----------------------------------------------------


def bn_bwd_reduce(dA: Tensor, Y: Tensor, scale: Optional[Tensor], shift: Optional[Tensor]) -> Tensor:
    dA, Y = _rows(dA), _rows(Y)
    M, C_ = Y.shape
    sums = torch.empty(2 * C_, dtype=torch.float64, device=Y.device)
    call("p2c_bn_bwd_reduce", ptr(dA), dA.stride(0), ptr(Y), Y.stride(0), ptr(scale), ptr(shift), M, C_, ptr(sums),
         stream_ptr())
    return sums


def pool_bwd_reduce(dOut: Tensor, Ymax: Tensor, Ymin: Tensor, scale: Tensor, shift: Tensor) -> Tensor:
    dOut = _rows(dOut)
    G, C_ = Ymax.shape
    sums = torch.empty(2 * C_, dtype=torch.float64, device=Ymax.device)
    call("p2c_pool_bwd_reduce", ptr(dOut), dOut.stride(0), ptr(Ymax), ptr(Ymin), ptr(scale), ptr(shift), G, C_, ptr(sums),
         stream_ptr())
    return sums


def bn_bwd_coef(sums: Tensor, count: int, gamma: Tensor, mean: Tensor, invstd: Tensor, training: bool,
                dgamma: Optional[Tensor], dbeta: Optional[Tensor]) -> Tensor:
    C_ = mean.shape[0]
    coef = torch.empty(3 * C_, dtype=torch.float32, device=mean.device)
    call("p2c_bn_bwd_coef", ptr(sums), count, ptr(gamma), ptr(mean), ptr(invstd), 1 if training else 0, ptr(coef),
         ptr(dgamma), ptr(dbeta), C_, stream_ptr())
    return coef


def bn_bwd_apply(dA: Tensor, Y: Tensor, scale: Tensor, shift: Tensor, coef: Tensor, out: Optional[Tensor] = None):
    dA, Y = _rows(dA), _rows(Y)
    M, C_ = Y.shape
    if out is None:
        out = torch.empty(M, C_, dtype=torch.float32, device=Y.device)
    call("p2c_bn_bwd_apply", ptr(dA), dA.stride(0), ptr(Y), Y.stride(0), ptr(scale), ptr(shift), ptr(coef), M, C_,
         ptr(out), out.stride(0), stream_ptr())
    return out


def pool_bwd_apply(dOut: Tensor, Ymax: Tensor, Ymin: Tensor, Y: Tensor, scale: Tensor, shift: Tensor, coef: Tensor,
                   group: int) -> Tensor:
    dOut, Y = _rows(dOut), _rows(Y)
    G, C_ = Ymax.shape
    dY = torch.empty(G * group, C_, dtype=torch.float32, device=Y.device)
    call("p2c_pool_bwd_apply", ptr(dOut), dOut.stride(0), ptr(Ymax), ptr(Ymin), ptr(Y), Y.stride(0), ptr(scale),
         ptr(shift), ptr(coef), G, group, C_, ptr(dY), dY.stride(0), stream_ptr())
    return dY


def bn_bwd_fusable(C_: int) -> bool:
    """p2c_bn_bwd_apply_fused keeps a thread on one group of four channels: C / 4 divides 256 (or is a multiple)."""
    return C_ % 4 == 0 and (256 % (C_ // 4) == 0 or (C_ // 4) % 256 == 0)


def bn_bwd_apply_fused(dA: Tensor, Y: Tensor, scale: Tensor, shift: Tensor, sums: Tensor, count: int, gamma: Tensor,
                       mean: Tensor, invstd: Tensor, training: bool, dgamma: Optional[Tensor],
                       dbeta: Optional[Tensor], out: Optional[Tensor] = None) -> Tensor:
    """bn_bwd_coef + bn_bwd_apply in one launch (the coefficients are evaluated by the applying kernel)."""
    dA, Y = _rows(dA), _rows(Y)
    M, C_ = Y.shape
    if out is None:
        out = torch.empty(M, C_, dtype=torch.float32, device=Y.device)
    call("p2c_bn_bwd_apply_fused", ptr(dA), dA.stride(0), ptr(Y), Y.stride(0), ptr(scale), ptr(shift), ptr(sums), count,
         ptr(gamma), ptr(mean), ptr(invstd), 1 if training else 0, ptr(dgamma), ptr(dbeta), M, C_, ptr(out),
         out.stride(0), stream_ptr())
    return out


def pool_bwd_apply_fused(dOut: Tensor, Ymax: Tensor, Ymin: Tensor, Y: Tensor, scale: Tensor, shift: Tensor,
                         sums: Tensor, count: int, gamma: Tensor, mean: Tensor, invstd: Tensor, training: bool,
                         dgamma: Optional[Tensor], dbeta: Optional[Tensor], group: int) -> Tensor:
    dOut, Y = _rows(dOut), _rows(Y)
    G, C_ = Ymax.shape
    dY = torch.empty(G * group, C_, dtype=torch.float32, device=Y.device)
    call("p2c_pool_bwd_apply_fused", ptr(dOut), dOut.stride(0), ptr(Ymax), ptr(Ymin), ptr(Y), Y.stride(0), ptr(scale),
         ptr(shift), ptr(sums), count, ptr(gamma), ptr(mean), ptr(invstd), 1 if training else 0, ptr(dgamma), ptr(dbeta),
         G, group, C_, ptr(dY), dY.stride(0), stream_ptr())
    return dY


def wgrad(dY: Tensor, X: Tensor, K: int, dW: Tensor, db: Optional[Tensor], in_scale: Optional[Tensor] = None,
          in_shift: Optional[Tensor] = None, mask_cf: Optional[Tensor] = None,
          precision: int = _lib.PREC_3XTF32) -> None:
    """dW (N, >=K rows of stride dW.stride(0)) += dY^T f(X);  db (N) += column sums of dY."""
    dY, X = _rows(dY), _rows(X)
    M, N = dY.shape
    dW2 = dW if dW.dim() == 2 else dW.reshape(dW.shape[0], -1)
    if dW2.stride(1) != 1 or dW2.shape[0] != N or dW2.shape[1] < K:
        raise _lib.P2CError(f"wgrad: dW {tuple(dW.shape)} does not match N={N}, K={K}")
    call("p2c_wgrad", ptr(dY), dY.stride(0), ptr(X), X.stride(0), ptr(in_scale), ptr(in_shift), ptr(mask_cf),
         0 if mask_cf is None else mask_cf.shape[2], M, N, K, ptr(dW2), dW2.stride(0), ptr(db), precision,
         stream_ptr())


def wgrad_on_tensor_cores(dY: Tensor, X: Tensor, K: int, has_mask: bool = False) -> bool:
    return bool(_lib.load().p2c_wgrad_path(dY.stride(0), X.stride(0), dY.shape[0], dY.shape[1], K, int(has_mask))) \
        and dY.data_ptr() % 16 == 0 and X.data_ptr() % 16 == 0


def sa_first_bwd(dY: Tensor, xyz: Tensor, new_xyz: Tensor, idx: Tensor, dQf: Optional[Tensor], dW: Tensor,
                 dbias: Optional[Tensor]) -> None:
    dY = _rows(dY)
    B, N, _ = xyz.shape
    S, ns = idx.shape[1], idx.shape[2]
    dW2 = dW.reshape(dW.shape[0], -1)
    call("p2c_sa_first_bwd", ptr(dY), dY.stride(0), ptr(xyz), ptr(new_xyz), ptr(idx), B, N, S, ns, dY.shape[1], ptr(dQf),
         0 if dQf is None else dQf.stride(0), ptr(dW2), dW2.stride(0), ptr(dbias), stream_ptr())


def group_bwd(dRows: Tensor, idx: Tensor, B: int, N: int, D: int) -> Tensor:
    """Backward of `group` w.r.t. the features: (B*N, D) scatter-add of the feature columns of the grouped rows."""
    dRows = _rows(dRows)
    idx = idx.contiguous()
    S, ns = idx.shape[1], idx.shape[2]
    dF = torch.zeros(B * N, D, dtype=torch.float32, device=dRows.device)
    call("p2c_group_bwd", ptr(dRows), dRows.stride(0), ptr(idx), B, N, S, ns, D, ptr(dF), dF.stride(0), stream_ptr())
    return dF


def three_nn_interp_bwd(dInterp: Tensor, idx: Optional[Tensor], w: Optional[Tensor], B: int, N: int, S: int) -> Tensor:
    dInterp = _rows(dInterp)
    D = dInterp.shape[1]
    dF = torch.empty(B * S, D, dtype=torch.float32, device=dInterp.device)
    call("p2c_three_nn_interp_bwd", ptr(dInterp), dInterp.stride(0), ptr(idx), ptr(w), B, N, S, D, ptr(dF), dF.stride(0),
         stream_ptr())
    return dF


def head_bwd(dOut: Tensor, mask_cf: Optional[Tensor], W: Tensor, B: int, N: int, H: Optional[Tensor] = None,
             scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, seed: Optional[Tensor] = None):
    """-> dA (B*N, C) [, A (B*N, C) = relu(bn(H)) * mask when H is given: the heads' input for their wgrad]."""
    dOut = _rows(dOut)
    W2 = W.reshape(W.shape[0], -1).contiguous()
    Nout, C_ = W2.shape
    dA = torch.empty(B * N, C_, dtype=torch.float32, device=dOut.device)
    A = None
    if H is not None:
        H = _rows(H)
        A = torch.empty(B * N, C_, dtype=torch.float32, device=dOut.device)
    call("p2c_head_bwd", ptr(dOut), dOut.stride(0), ptr(mask_cf), ptr(_seed(seed)), ptr(W2), B, N, C_, Nout, ptr(dA), dA.stride(0),
         ptr(H), 0 if H is None else H.stride(0), ptr(scale), ptr(shift), ptr(A), 0 if A is None else A.stride(0),
         stream_ptr())
    return dA if H is None else (dA, A)


def adam_step(params: Tensor, grads: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, lr: float, step: int,
              betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, grad_scale: float = 1.0) -> None:
    """In-place Adam on flat fp32 buffers (torch.optim.Adam semantics, no amsgrad)."""
    need_cuda(params, grads, exp_avg, exp_avg_sq)
    n = params.numel()
    for t in (params, grads, exp_avg, exp_avg_sq):
        if t.numel() != n or not t.is_contiguous() or t.dtype != torch.float32:
            raise _lib.P2CError("adam_step: flat contiguous float32 buffers of equal length expected")
    call("p2c_adam_step", ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), n, float(lr), float(betas[0]),
         float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale), stream_ptr())
