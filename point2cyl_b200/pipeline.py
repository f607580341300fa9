"""Forward + loss of Point2Cyl as a sequence of libp2c kernels (point-major, no permutes).

This is the engine behind the drop-in modules (point2cyl_b200/dropin/): `backbone_forward` is what
`backbone.forward` runs (models/pointnet_extrusion.py:37-66 in the reference) and `loss_forward`
is the loss half of a training step (train_Point2Cyl_without_sketch.py:246-353).  Activations are
2-D row matrices (B*points, C); every BatchNorm+ReLU is folded into the next layer's operand load
and every layer's statistics / max-pool into its epilogue (see include/point2cyl.h: p2c_linear).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib, ops

Tensor = torch.Tensor

_PRECISIONS = {"fp32": _lib.PREC_FP32, "3xtf32": _lib.PREC_3XTF32, "bf16": _lib.PREC_BF16}
_default_precision = "3xtf32"   # tcgen05, fp32-faithful; layers it does not take run the fp32 SIMT kernel


def set_precision(name: str) -> None:
    """'fp32' (SIMT), '3xtf32' (tcgen05, fp32-faithful; default) or 'bf16' (tcgen05 kind::f16 on bf16 operands with
    fp32 accumulation - BASELINE.json's bf16 MLP-stack configuration, NOT within the 1e-4 parity bar)."""
    global _default_precision
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
    _default_precision = name


def get_precision() -> str:
    return _default_precision


# Dropout of the head (models/pointnet_extrusion.py:60, F.dropout(p=0.5), always on).  Default: the head kernels draw
# the mask themselves (Philox4x32-10 keyed by two int64 words from torch's CUDA generator - torch.manual_seed controls
# it, CUDA-graph replays get fresh words), so no (B,128,N) mask tensor is written or read.  `dropout_mask_fn`, when
# set, is called like F.dropout(ones(B,128,N), p=0.5) and must return the multiplicative mask: tests inject fixed
# masks through it, and `dropout_mask_fn = torch_dropout_mask` reproduces the reference's own mask stream for a
# shared seed (same torch op on the same shape) at the cost of 268 MB of traffic per step at B=32 x N=8192.
dropout_mask_fn = None

# fp3 (one source point per cloud) through the linearity of its first conv (feature_propagation -> ops.linear_group_bias)
group_bias_enabled = True

# sa1 without its first layer's activations (set_abstraction -> ops.sa_xyz_linear); False = always materialise them
xyz_first_enabled = True

# eval-mode BatchNorm: a feature-less SA level (sa1) as ONE kernel - gather, three conv/BN/ReLU layers, max-pool
# (set_abstraction -> ops.sa_stack_fused); False = the per-layer kernels
stack_fused_enabled = True


def torch_dropout_mask(ones: Tensor, p: float = 0.5) -> Tensor:
    return F.dropout(ones, p=p)


def draw_dropout_seed(device) -> Tensor:
    return torch.randint(-2 ** 62, 2 ** 62, (2,), dtype=torch.long, device=device)


def _bn_momentum(bn) -> float:
    if bn.momentum is None:  # cumulative moving average
        return 1.0 / float(int(bn.num_batches_tracked) + 1)
    return float(bn.momentum)


@dataclass
class Affine:
    """A pending BatchNorm+ReLU: y = max(x*scale + shift, 0), applied by whoever reads x next.  `bn` (ops.PendingBN)
    is set while even the computation of scale / shift is still pending: the first consumer kernel folds it in its
    prologue (p2c_bn_fold) and writes the two arrays; `resolved()` forces that with the stand-alone kernel."""
    scale: Tensor
    shift: Tensor
    bn: Optional["ops.PendingBN"] = None

    def resolved(self) -> "Affine":
        if self.bn is not None:
            self.bn.resolve()
        return self


class _ForwardScratch:
    """Per-forward batching of what used to be one tiny launch per layer: ONE zeroed float64 buffer for all BatchNorm
    sum / sum-of-squares accumulators, ONE p2c_split_tf32_multi launch for every streamed-weight layer, ONE
    _foreach_add_ for the num_batches_tracked counters.  mlp_stack falls back to per-layer calls without it."""

    def __init__(self, net, precision: Optional[str]):
        prec = _PRECISIONS[precision or _default_precision]
        dev = next(net.parameters()).device
        bns = [m for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        self.stats = torch.zeros(2 * sum(m.num_features for m in bns), dtype=torch.float64, device=dev)
        self._stats_off = 0
        self.tracked = []
        self.wsplit = {}
        if prec == _lib.PREC_3XTF32:
            convs = [m for m in net.modules() if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv2d))]
            # column slices registered by feature_propagation's linearity path (split next to the full matrices)
            slices = getattr(net, "_p2c_split_slices", {})
            big = []          # (dict key, weight view)
            for c in convs:
                N_, K_ = c.weight.shape[0], c.weight[0].numel()
                # the layers p2c_linear sends to the streamed-weight kernel when given a split copy (pure function of
                # the shape; row stride pad4(K) as mlp_stack / the concat buffers produce it)
                if _lib.load().p2c_linear_path(ops.pad4(K_), 1 << 20, N_, K_, 0, 0, prec, 1) == 2:
                    big.append((id(c), c.weight))
                ks = slices.get(id(c))
                if ks:
                    big.append(((id(c), ks), c.weight.reshape(N_, -1)[:, :ks]))
            if big:
                cache = getattr(net, "_p2c_wsplit_buffers", None)
                key = tuple((k, w.data_ptr()) for k, w in big)
                if cache is None or cache[0] != key:
                    cache = (key, None)
                outs = ops.split_tf32_multi([w for _, w in big], cache[1])
                net._p2c_wsplit_buffers = (key, outs)
                self.wsplit = {k: o for (k, _), o in zip(big, outs)}

        self.floats = torch.empty(4 * sum(m.num_features for m in bns), dtype=torch.float32, device=dev)
        self._floats_off = 0

    def take_floats(self, n: int) -> Tensor:
        out = self.floats[self._floats_off:self._floats_off + n]
        self._floats_off += n
        return out if out.numel() == n else None

    def take_stats(self, n: int) -> Tensor:
        out = self.stats[self._stats_off:self._stats_off + 2 * n]
        self._stats_off += 2 * n
        if out.numel() != 2 * n:
            raise _lib.P2CError("forward scratch: more BatchNorm layers ran than the network declares")
        return out

    def finish(self):
        if self.tracked:
            torch._foreach_add_(self.tracked, 1)


_scratch: Optional[_ForwardScratch] = None


def mlp_stack(X: Tensor, K: int, convs, bns, training: bool, pool_group: int = 0,
              in_affine: Optional[Affine] = None, in_mask: Optional[Tensor] = None,
              precision: Optional[str] = None, tag: str = "mlp", first_layer=None, tape: Optional[list] = None,
              xyz_first: Optional[dict] = None):
    """Runs [conv1x1 -> BN -> ReLU] * L over the rows of X.

    Returns (Y_last_raw, Affine_last) — or, with pool_group, the pooled post-BN/ReLU features
    (rows/pool_group, C_last).  The BN of layer i is applied inside layer i+1's operand load.
    tape: when given, every layer appends what point2cyl_b200.backward needs (raw input/output, folded BN,
    batch mean / invstd, pooled extrema) and pooled layers keep their raw output.
    xyz_first (dict xyz, new_xyz, gidx[, mom]): a feature-less SA level whose first conv is NOT materialised - its
    BatchNorm statistics come in closed form (ops.sa_xyz_stats) and the second layer recomputes its rows in the operand
    transform (ops.sa_xyz_linear); forward only (no tape).
    """
    prec = _PRECISIONS[precision or _default_precision]
    M = X.shape[0] if X is not None else None
    aff = in_affine
    mask = in_mask
    last = len(convs) - 1
    save = tape is not None
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        N = conv.weight.shape[0]
        use_batch_stats = training or (bn.running_mean is None)
        stats = None
        if use_batch_stats:
            stats = _scratch.take_stats(N) if _scratch is not None else \
                torch.zeros(2 * N, dtype=torch.float64, device=conv.weight.device)
        pool = pool_group if i == last else 0
        _lib.set_tag(f"{tag}.{i}")
        fused_first = i == 0 and first_layer is not None
        Ymax = Ymin = None
        if xyz_first is not None and i == 0:
            g = xyz_first
            M = g["gidx"].numel()
            if use_batch_stats and g.get("mom") is None:
                g["mom"] = ops.group_moments(g["xyz"], g["new_xyz"], g["gidx"])
            Y = None        # its statistics: derived from the moments inside the next layer's prologue
        elif xyz_first is not None and i == 1:
            g = xyz_first
            res = ops.sa_xyz_linear(g["xyz"], g["new_xyz"], g["gidx"], convs[0].weight, convs[0].bias, conv.weight,
                                    conv.bias, scale0=aff.scale, shift0=aff.shift, bn0=aff.bn, stats=stats,
                                    pool_group=pool, want_y=(pool == 0), moments=g.get("mom"))
            if res is None:
                raise _lib.P2CError("sa_xyz_linear: shape not taken by the tensor-core kernel")
            if pool:
                Y, Ymax, Ymin = res
            else:
                Y = res
        elif fused_first:
            # the level's first conv comes fused with the grouping gather (p2c_sa_first_layer)
            Y = first_layer(stats)
            M = Y.shape[0]
        else:
            wsplit = _scratch.wsplit.get(id(conv)) if _scratch is not None else None
            if wsplit is None or not ops.needs_split(X, N, K, mask is not None, pool, prec):
                wsplit = ops.weight_operand(X, conv.weight, N, K, mask is not None, pool, prec)
            res = ops.linear(X, conv.weight, conv.bias, K=K, w_split=wsplit,
                             in_scale=None if aff is None else aff.scale,
                             in_shift=None if aff is None else aff.shift,
                             in_bn=None if aff is None else aff.bn,
                             in_mask=mask, stats=stats, pool_group=pool, want_y=(pool == 0 or save), precision=prec)
            if pool:
                Y, Ymax, Ymin = res
            else:
                Y = res
        # the finalisation itself is deferred: the kernel that reads this layer's output folds it (p2c_bn_fold)
        pend = ops.PendingBN(stats, M, bn.weight, bn.bias, bn.eps, _bn_momentum(bn), use_batch_stats,
                             bn.running_mean, bn.running_var, save=save,
                             out=None if _scratch is None else _scratch.take_floats(4 * N))
        scale, shift = pend.scale, pend.shift
        if training and bn.num_batches_tracked is not None:
            if _scratch is not None:
                _scratch.tracked.append(bn.num_batches_tracked)
            else:
                bn.num_batches_tracked.add_(1)
        if save:
            tape.append(dict(X=None if fused_first else X, K=K, in_aff=aff, in_mask=mask, conv=conv, bn=bn, Y=Y,
                             scale=scale, shift=shift, mean=pend.mean, invstd=pend.invstd, M=M, pool=pool, Ymax=Ymax, Ymin=Ymin,
                             batch_stats=use_batch_stats, fused_first=fused_first))
        aff = Affine(scale, shift, pend)
        mask = None
        if pool:
            return ops.pool_bn_relu(Ymax, Ymin, scale, shift, bn=pend)
        X = Y
        K = N
    return X, aff


def set_abstraction(sa, xyz: Tensor, feats: Optional[Tensor], start: Optional[Tensor],
                    trace: Optional[dict] = None, precision: Optional[str] = None, tag: str = "sa",
                    fused_first: bool = True, tape: Optional[dict] = None, geo=None, mom: Optional[Tensor] = None):
    """PointNetSetAbstraction in point-major form (models/pointnet_util.py:181-207).
    xyz (B,N,3), feats (B*N, D) rows or None -> (new_xyz (B,S,3), new_feats (B*S, C)).
    tape: dict filled with what the backward of this level needs.
    geo: (fps_idx, new_xyz, group_idx) computed earlier by `geometry_forward` (they depend on xyz only);
    mom: ops.group_moments of the same grouping, when the geometry stage already reduced them."""
    B, N, _ = xyz.shape
    D = 0 if feats is None else feats.shape[1]
    _lib.set_tag(tag)
    layers = None
    if tape is not None:
        layers = []
        tape.update(kind="sa", module=sa, xyz=xyz, feats=feats, B=B, N=N, D=D, layers=layers, fused=False)
    if sa.group_all:
        rows = ops.group(xyz, feats, None, None)
        new_xyz = torch.zeros(B, 1, 3, dtype=torch.float32, device=xyz.device)
        pool = N
    else:
        if geo is not None:
            fps_idx, new_xyz, gidx = geo
        else:
            fps_idx, new_xyz = ops.fps(xyz, sa.npoint, start)
            gidx = ops.ball_query(sa.radius, sa.nsample, xyz, new_xyz)
        pool = sa.nsample
        if trace is not None:
            trace["fps_idx"], trace["group_idx"] = fps_idx, gidx
        conv0 = sa.mlp_convs[0]
        C0 = conv0.weight.shape[0]
        prec = _PRECISIONS[precision or _default_precision]
        if C0 in (64, 128) and len(sa.mlp_convs) > 1 and fused_first:
            # linearity of the 1x1 conv: the feature half of the first layer is computed once per source point
            W0 = conv0.weight.reshape(C0, -1)
            Qf = None
            if feats is not None:
                _lib.set_tag(tag + ".q")
                Wf = W0[:, 3:].contiguous()
                Qf = ops.linear(feats, Wf, None, K=D, precision=prec,
                                w_split=ops.weight_operand(feats, Wf, C0, D, False, 0, prec))

            def first_layer(stats):
                return ops.sa_first_layer(xyz, new_xyz, gidx, Qf, W0, conv0.bias, stats)

            rows = B * gidx.shape[1] * gidx.shape[2]
            N1 = sa.mlp_convs[1].weight.shape[0]
            pool1 = pool if len(sa.mlp_convs) == 2 else 0
            if (stack_fused_enabled and feats is None and tape is None and prec == _lib.PREC_3XTF32 and
                    not sa.training and len(sa.mlp_convs) == 3 and
                    all(bn.running_mean is not None for bn in sa.mlp_bns)):
                # running statistics: every BatchNorm of the level is known before the launch, so the whole level is one
                # kernel and none of its activations reaches HBM (csrc/sa_stack_tc.cu)
                pend = [ops.PendingBN(None, rows, bn.weight, bn.bias, bn.eps, _bn_momentum(bn), False, bn.running_mean,
                                      bn.running_var,
                                      out=None if _scratch is None else _scratch.take_floats(4 * bn.weight.shape[0]))
                        for bn in sa.mlp_bns]
                _lib.set_tag(tag + ".fused")
                out = ops.sa_stack_fused(xyz, new_xyz, gidx, list(sa.mlp_convs), pend)
                if out is not None:
                    return new_xyz, out
            if (xyz_first_enabled and feats is None and tape is None and prec == _lib.PREC_3XTF32 and
                    rows % max(pool1, 1) == 0 and
                    _lib.load().p2c_linear_path(ops.pad4(C0), rows, N1, C0, 0, pool1, prec, 0) == 1):
                # no input features and no tape: the first conv's rows are never written - closed-form BatchNorm
                # statistics + recomputation inside the second layer's operand transform (ops.sa_xyz_linear)
                out = mlp_stack(None, 3, sa.mlp_convs, sa.mlp_bns, sa.training, pool_group=pool, precision=precision,
                                tag=tag, xyz_first=dict(xyz=xyz, new_xyz=new_xyz, gidx=gidx, mom=mom))
                return new_xyz, out

            if tape is not None:
                tape.update(fused=True, new_xyz=new_xyz, gidx=gidx)
            out = mlp_stack(None, 3 + D, sa.mlp_convs, sa.mlp_bns, sa.training, pool_group=pool, precision=precision,
                            tag=tag, first_layer=first_layer, tape=layers)
            return new_xyz, out
        if tape is not None:
            tape.update(gidx=gidx, grouped=True)     # un-fused path: the gather's backward is p2c_group_bwd
        rows = ops.group(xyz, feats, new_xyz, gidx)
    out = mlp_stack(rows, 3 + D, sa.mlp_convs, sa.mlp_bns, sa.training, pool_group=pool, precision=precision,
                    tag=tag, tape=layers)
    return new_xyz, out


def feature_propagation(fp, xyz1: Tensor, xyz2: Tensor, feats1: Optional[Tensor], feats2: Tensor,
                        materialize: bool = True, precision: Optional[str] = None, tag: str = "fp",
                        tape: Optional[dict] = None, nn=None):
    """PointNetFeaturePropagation in point-major form (models/pointnet_util.py:283-320).
    feats1 (B*N, D1) or None, feats2 (B*S, D2) -> (B*N, C) post-BN/ReLU rows (materialize=True) or
    (raw rows, Affine) for a consumer that folds the last BN+ReLU into its own load."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    D1 = 0 if feats1 is None else feats1.shape[1]
    D2 = feats2.shape[1]
    _lib.set_tag(tag)
    prec = _PRECISIONS[precision or _default_precision]
    conv0 = fp.mlp_convs[0]
    C0 = conv0.weight.shape[0]
    if (S == 1 and feats1 is not None and tape is None and group_bias_enabled and prec == _lib.PREC_3XTF32 and
            N % 32 == 0 and D2 <= 3072 and B <= 256 and feats1.stride(0) % 4 == 0 and
            _lib.load().p2c_linear_path(feats1.stride(0), B * N, C0, D1, 0, 0, prec, 1) == 2):
        # one source point per cloud (fp3): `points2.repeat(1, N, 1)` + concat + conv (:298-299, :312, :317) is, by
        # linearity, feats1 W[:, :D1]^T + (feats2 W[:, D1:]^T + b)[cloud] - the broadcast rows and the (B*N, D1+D2)
        # concat buffer never exist, and the layer's K drops from D1 + D2 to D1 (fp3.0: 1280 -> 256)
        W0 = conv0.weight.reshape(C0, -1)

        def first_layer(stats):
            per_cloud = ops.linear_small(feats2, W0[:, D1:], conv0.bias)            # (B, C0), fp32
            wsplit = _scratch.wsplit.get((id(conv0), D1)) if _scratch is not None else None
            if wsplit is None:
                wsplit = ops.split_tf32_multi([W0[:, :D1]])[0]
                net_slices = getattr(fp, "_p2c_owner_slices", None)
                if net_slices is not None:
                    net_slices[id(conv0)] = D1      # later forwards split the slice with all the other weights
            return ops.linear_group_bias(feats1, wsplit, per_cloud, N, C0, D1, stats=stats)

        Y, aff = mlp_stack(None, D1 + D2, fp.mlp_convs, fp.mlp_bns, fp.training, precision=precision, tag=tag,
                           first_layer=first_layer)
        _lib.set_tag(tag)
        if materialize:
            return ops.bn_relu_apply(Y, aff.scale, aff.shift, bn=aff.bn)
        return Y, aff
    buf = torch.empty(B * N, D1 + D2, dtype=torch.float32, device=xyz1.device)
    if feats1 is not None:
        buf[:, :D1].copy_(feats1)           # skip features first (:312)
    layers = None
    nn_idx = nn_w = None
    if nn is not None and S > 1:           # neighbours found earlier by `geometry_forward`: gather only
        nn_idx, nn_w = nn
        ops.three_nn_gather(feats2, nn_idx, nn_w, S, out=buf[:, D1:])
    elif tape is not None:
        _, nn_idx, nn_w = ops.three_nn_interp(xyz1, xyz2, feats2, out=buf[:, D1:], want_idx=True)
    else:
        ops.three_nn_interp(xyz1, xyz2, feats2, out=buf[:, D1:])
    if tape is not None:
        layers = []
        tape.update(kind="fp", module=fp, B=B, N=N, S=S, D1=D1, D2=D2, nn_idx=nn_idx, nn_w=nn_w, layers=layers)
    Y, aff = mlp_stack(buf, D1 + D2, fp.mlp_convs, fp.mlp_bns, fp.training, precision=precision, tag=tag, tape=layers)
    _lib.set_tag(tag)
    if materialize:
        return ops.bn_relu_apply(Y, aff.scale, aff.shift, bn=aff.bn)
    return Y, aff


def draw_fps_start(B: int, N: int, device) -> Tensor:
    """First FPS centroid, drawn exactly like the reference: CPU generator, then moved
    (models/pointnet_util.py:75)."""
    return torch.randint(0, N, (B,), dtype=torch.long).to(device)


@dataclass
class Geometry:
    """Everything of a backbone forward that depends on the point COORDINATES only (no weights, no features): the
    two FPS levels, both ball queries and the 3-NN neighbours / weights of fp1 and fp2 (fp3 interpolates from a single
    point).  A pipelined caller computes it for batch i+1 on a second stream while batch i runs through the per-point
    MLP layers (graph.PipelinedForwardLoss); `backbone_forward` computes it inline."""
    xyz: Tensor
    fps1: Tensor
    l1_xyz: Tensor
    gidx1: Tensor
    fps2: Tensor
    l2_xyz: Tensor
    gidx2: Tensor
    nn1_idx: Tensor
    nn1_w: Tensor
    nn2_idx: Tensor
    nn2_w: Tensor
    mom1: Optional[Tensor] = None      # ops.group_moments of level 1 (closed-form BatchNorm statistics of sa1.0)

    @staticmethod
    def empty(net, B: int, N: int, device) -> "Geometry":
        S1, S2, ns1, ns2 = net.sa1.npoint, net.sa2.npoint, net.sa1.nsample, net.sa2.nsample
        e = lambda shape, dt: torch.empty(*shape, dtype=dt, device=device)
        return Geometry(e((B, N, 3), torch.float32), e((B, S1), torch.long), e((B, S1, 3), torch.float32),
                        e((B, S1, ns1), torch.long), e((B, S2), torch.long), e((B, S2, 3), torch.float32),
                        e((B, S2, ns2), torch.long), e((B, N, 3), torch.long), e((B, N, 3), torch.float32),
                        e((B, S1, 3), torch.long), e((B, S1, 3), torch.float32),
                        torch.zeros(int(_lib.load().p2c_group_moments_size()), dtype=torch.float64, device=device))


def geometry_forward(net, xyz: Tensor, fps_start: Optional[Sequence[Tensor]] = None,
                     out: Optional[Geometry] = None, moments: bool = True) -> Geometry:
    """FPS -> ball query (both levels) and the 3-NN searches of fp1 / fp2 for coordinates xyz (B,N,3) contiguous:
    models/pointnet_util.py:63-107 and :301-305.  out: a Geometry whose buffers are overwritten (static across
    CUDA-graph replays)."""
    _lib.need_cuda(xyz)
    B, N, _ = xyz.shape
    dev = xyz.device
    g = out
    _lib.set_tag("sa1")
    s1 = fps_start[0] if fps_start is not None else draw_fps_start(B, N, dev)
    fps1, l1_xyz = ops.fps(xyz, net.sa1.npoint, s1, out=None if g is None else (g.fps1, g.l1_xyz))
    gidx1 = ops.ball_query(net.sa1.radius, net.sa1.nsample, xyz, l1_xyz, out=None if g is None else g.gidx1)
    mom1 = ops.group_moments(xyz, l1_xyz, gidx1, out=None if g is None else g.mom1) if moments else None
    _lib.set_tag("sa2")
    s2 = fps_start[1] if fps_start is not None else draw_fps_start(B, l1_xyz.shape[1], dev)
    fps2, l2_xyz = ops.fps(l1_xyz, net.sa2.npoint, s2, out=None if g is None else (g.fps2, g.l2_xyz))
    gidx2 = ops.ball_query(net.sa2.radius, net.sa2.nsample, l1_xyz, l2_xyz, out=None if g is None else g.gidx2)
    _lib.set_tag("fp1")
    nn1 = ops.three_nn_search(xyz, l1_xyz, out=None if g is None else (g.nn1_idx, g.nn1_w))
    _lib.set_tag("fp2")
    nn2 = ops.three_nn_search(l1_xyz, l2_xyz, out=None if g is None else (g.nn2_idx, g.nn2_w))
    if g is not None:
        if g.xyz.data_ptr() != xyz.data_ptr():
            g.xyz.copy_(xyz)
        return g
    return Geometry(xyz, fps1, l1_xyz, gidx1, fps2, l2_xyz, gidx2, nn1[0], nn1[1], nn2[0], nn2[1], mom1)


def backbone_forward(net, x: Tensor, fps_start: Optional[Sequence[Tensor]] = None,
                     trace: Optional[dict] = None, precision: Optional[str] = None,
                     tape: Optional[dict] = None, geo: Optional[Geometry] = None) -> List[Tensor]:
    """models/pointnet_extrusion.py:37-66.  x (B,N,3[+3]) -> [ (B,N,o_i) ] (views of one buffer).
    tape: dict that receives the per-stage records point2cyl_b200.backward.backbone_backward consumes.
    geo: the coordinate-only stage computed ahead of time (`geometry_forward`); None = compute it here."""
    _lib.need_cuda(x)
    B, N, Cx = x.shape
    x = x.float()
    xyz = x[:, :, :3].contiguous()
    feats0 = x[:, :, 3:].reshape(B * N, Cx - 3).contiguous() if Cx > 3 else None
    dev = x.device
    rec = (lambda: {}) if tape is not None else (lambda: None)
    r_sa1, r_sa2, r_sa3, r_fp3, r_fp2, r_fp1 = rec(), rec(), rec(), rec(), rec(), rec()
    if geo is None:
        geo = geometry_forward(net, xyz, fps_start,
                               moments=tape is None and feats0 is None and xyz_first_enabled and net.training)
    global _scratch
    outer, _scratch = _scratch, _ForwardScratch(net, precision)
    try:
        return _backbone_features(net, xyz, feats0, geo, trace, precision, tape, r_sa1, r_sa2, r_sa3, r_fp3, r_fp2,
                                  r_fp1, B, N, dev)
    finally:
        _scratch.finish()
        _scratch = outer


def _backbone_features(net, xyz, feats0, geo, trace, precision, tape, r_sa1, r_sa2, r_sa3, r_fp3, r_fp2, r_fp1, B, N,
                       dev):
    """The feature stage of backbone_forward (everything that involves weights)."""
    t1 = {} if trace is not None else None
    l1_xyz, l1 = set_abstraction(net.sa1, xyz, feats0, None, t1, precision, tag="sa1", tape=r_sa1,
                                 geo=(geo.fps1, geo.l1_xyz, geo.gidx1), mom=geo.mom1)
    t2 = {} if trace is not None else None
    l2_xyz, l2 = set_abstraction(net.sa2, l1_xyz, l1, None, t2, precision, tag="sa2", tape=r_sa2,
                                 geo=(geo.fps2, geo.l2_xyz, geo.gidx2))
    l3_xyz, l3 = set_abstraction(net.sa3, l2_xyz, l2, None, None, precision, tag="sa3", tape=r_sa3)
    if not hasattr(net, "_p2c_split_slices"):
        net._p2c_split_slices = {}
    net.fp3._p2c_owner_slices = net._p2c_split_slices
    l4 = feature_propagation(net.fp3, l2_xyz, l3_xyz, l2, l3, precision=precision, tag="fp3", tape=r_fp3)
    l5 = feature_propagation(net.fp2, l1_xyz, l2_xyz, l1, l4, precision=precision, tag="fp2", tape=r_fp2,
                             nn=(geo.nn2_idx, geo.nn2_w))
    y6, aff6 = feature_propagation(net.fp1, xyz, l1_xyz, feats0, l5, materialize=False, precision=precision,
                                    tag="fp1", tape=r_fp1, nn=(geo.nn1_idx, geo.nn1_w))
    # FC head: fc1 -> bn1 -> ReLU -> dropout(p=.5, always on, :60) -> fc2 heads
    head_layers = [] if tape is not None else None
    h, aff_h = mlp_stack(y6, y6.shape[1], [net.fc1], [net.bn1], net.training, in_affine=aff6,
                         precision=precision, tag="fc1", tape=head_layers)
    mask_cf = seed = None
    if dropout_mask_fn is not None:
        # explicit mask in the reference's own (B,128,N) channel-first layout (consumed without a transpose copy)
        mask_cf = dropout_mask_fn(torch.ones(B, h.shape[1], N, dtype=torch.float32, device=dev), p=0.5)
        if tuple(mask_cf.shape) != (B, h.shape[1], N):
            raise _lib.P2CError("dropout mask must keep the (B,128,N) shape")
        mask_cf = mask_cf.contiguous()
    else:
        seed = draw_dropout_seed(dev)
    Wcat = torch.cat([fc.weight.reshape(fc.weight.shape[0], -1) for fc in net.fc2], dim=0)
    bcat = torch.cat([fc.bias for fc in net.fc2], dim=0)
    _lib.set_tag("fc2")
    out = ops.head_masked(h, aff_h.scale, aff_h.shift, mask_cf, Wcat, bcat, B, N, seed=seed, bn=aff_h.bn,
                          precision=_PRECISIONS[precision or _default_precision])
    if trace is not None:
        trace.update(sa1=t1, sa2=t2, l1_xyz=l1_xyz, l1=l1, l2_xyz=l2_xyz, l2=l2, l3=l3, l4=l4, l5=l5,
                     y6=y6, aff6=aff6, h=h, aff_h=aff_h)
    if tape is not None:
        tape.update(B=B, N=N, sa1=r_sa1, sa2=r_sa2, sa3=r_sa3, fp3=r_fp3, fp2=r_fp2, fp1=r_fp1,
                    head=dict(layers=head_layers, h=h, aff_h=aff_h, mask_cf=mask_cf, seed=seed, Wcat=Wcat, fc2=list(net.fc2)),
                    out=out, feats0=feats0)
    results, c0 = [], 0
    out3 = out.reshape(B, N, out.shape[1])
    for fc in net.fc2:
        o = fc.weight.shape[0]
        results.append(out3[:, :, c0:c0 + o])
        c0 += o
    return results


# ---- loss ----------------------------------------------------------------------------------------


def loss_forward(pcs: Tensor, X_raw: Tensor, W_raw: Tensor, gt_normals: Tensor, gt_inst: Tensor,
                 gt_bb: Tensor, gt_axes: Tensor, gt_centers: Tensor,
                 weights=(1.0, 1.0, 1.0, 1.0, 1.0), norm_eig: bool = False) -> Dict[str, Tensor]:
    """train_Point2Cyl_without_sketch.py:246-353 with all five --pred_* branches on.
    weights = (seg, normal, bb, extrusion, centre).  The assignment (losses.py:43, scipy on the host upstream) runs on
    the device (p2c_hungarian): no host sync anywhere in the step."""
    B, N, twoK = W_raw.shape
    K = twoK // 2
    _lib.set_tag("loss")
    from . import autograd as ag
    losses, match, n_gt, E_AX, centers, per_seg, per_cloud, stats = ag.fused_loss(
        X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers, weights, norm_eig)
    mask = torch.arange(K, device=pcs.device)[None, :] < n_gt[:, None]
    return dict(total=losses[0], normal=losses[1], miou=losses[2], bb=losses[3], axis=losses[4],
                center=losses[5], losses=losses, matching_indices=match, mask=mask, E_AX=E_AX,
                centers=centers, per_seg=per_seg, per_cloud=per_cloud, n_gt=n_gt, stats=stats)


def forward_loss(net, batch: Dict[str, Tensor], fps_start=None, weights=(1.0,) * 5,
                 norm_eig: bool = False, precision: Optional[str] = None, geo: Optional[Geometry] = None) -> Dict[str, Tensor]:
    """One forward+loss pass: the unit BASELINE.json's clouds/s metric counts."""
    X_raw, W_raw = backbone_forward(net, batch["pcs"], fps_start, precision=precision, geo=geo)
    out = loss_forward(batch["pcs"], X_raw, W_raw, batch["normals"], batch["inst"], batch["bb"],
                       batch["axes"], batch["centers"], weights, norm_eig)
    out.update(X_raw=X_raw, W_raw=W_raw)
    return out
