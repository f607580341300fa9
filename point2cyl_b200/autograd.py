"""torch.autograd.Function wrappers: each forward AND backward runs libp2c.so kernels (no torch-eager math on the
point dimension).  This is what lets the unmodified training scripts call `total_loss.backward()` on results of the
drop-in modules (train_Point2Cyl_without_sketch.py:367).

  SegfitStatsW     per-cloud sufficient statistics of soft assignments (function-level losses.py / data_utils.py)
  Eig3x3Smallest   eigenvector of the smallest eigenvalue (torch.symeig(...)[1][:, :, 0], data_utils.py:170-171)
  FusedLoss        the whole loss block of the training script on the network's raw outputs
  Backbone         the whole PointNet++ backbone as ONE autograd node (point2cyl_b200.backward does the work)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

Tensor = torch.Tensor


class SegfitStatsW(torch.autograd.Function):
    """stats (B, stride(K)) = p2c_segfit_stats_w(...); differentiable w.r.t. Wb, Wc and X."""

    @staticmethod
    def forward(ctx, Wb, Wc, X, normalize_x, pcs, gt_normals, inst, bb):
        ctx.save_for_backward(Wb, Wc, X, pcs, gt_normals, inst)
        ctx.normalize_x = bool(normalize_x)
        return ops.segfit_stats_w(Wb, Wc, X, normalize_x, pcs, gt_normals, inst, bb)

    @staticmethod
    def backward(ctx, dstats):
        Wb, Wc, X, pcs, gt_normals, inst = ctx.saved_tensors
        need_wb, need_wc, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dX, dWb, dWc = ops.segfit_backward_w(dstats, Wb, Wc, X, ctx.normalize_x, pcs, gt_normals, inst,
                                             want_dx=need_x, want_dwc=need_wc)
        return (dWb if need_wb else None, dWc if need_wc else None, dX if need_x else None, None, None, None, None,
                None)


def segfit_stats_w(Wb, Wc=None, X=None, normalize_x=False, pcs=None, gt_normals=None, inst=None, bb=None):
    return SegfitStatsW.apply(Wb, Wc, X, normalize_x, pcs, gt_normals, inst, bb)


class Eig3x3Smallest(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M):
        vec, ev = ops.eig3x3_smallest(M)
        ctx.save_for_backward(M)
        ctx.mark_non_differentiable(ev)
        return vec, ev

    @staticmethod
    def backward(ctx, gvec, _gev):
        (M,) = ctx.saved_tensors
        return ops.eig3x3_backward(M, gvec)


def eig3x3_smallest(M):
    return Eig3x3Smallest.apply(M)


class FusedLoss(torch.autograd.Function):
    """train_Point2Cyl_without_sketch.py:246-353 on (X_raw, W_raw): losses (6,) = {total, normal, miou, bb, axis,
    centre} plus non-differentiable by-products."""

    @staticmethod
    def forward(ctx, X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers, weights, norm_eig):
        B, N, twoK = W_raw.shape
        K = twoK // 2
        stats = ops.segfit_stats(X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, K)
        cost, n_gt = ops.segfit_cost(stats, K)
        match = ops.hungarian(cost, n_gt)
        bb_sum = ops.bb_loss_sums(W_raw, gt_bb, match, n_gt, K)
        losses, E_AX, centers, per_seg, per_cloud = ops.loss_finalize(
            stats, bb_sum, match, n_gt, gt_axes, gt_centers, N, K, norm_eig, weights)
        ctx.save_for_backward(X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers, stats, match, n_gt)
        ctx.weights, ctx.norm_eig, ctx.K, ctx.N = tuple(float(w) for w in weights), bool(norm_eig), K, N
        ctx.mark_non_differentiable(match, n_gt, E_AX, centers, per_seg, per_cloud, stats)
        return losses, match, n_gt, E_AX, centers, per_seg, per_cloud, stats

    @staticmethod
    def backward(ctx, g, *_):
        X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers, stats, match, n_gt = ctx.saved_tensors
        w = torch.tensor(ctx.weights, dtype=torch.float32, device=g.device)
        g = g.float()
        # losses = {total, normal, miou, bb, axis, centre}; weights = {seg, normal, bb, axis, centre}
        eff = w * g[0] + torch.stack([g[2], g[1], g[3], g[4], g[5]])
        dstats = ops.loss_backward_coef(stats, match, n_gt, gt_axes, gt_centers, eff, ctx.N, ctx.K, ctx.norm_eig)
        d_out = ops.segfit_backward(X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, dstats, match, n_gt, eff, ctx.K)
        B, N = gt_inst.shape
        d3 = d_out.reshape(B, N, -1)
        return (d3[:, :, :3], d3[:, :, 3:]) + (None,) * 8


def fused_loss(X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers, weights, norm_eig):
    return FusedLoss.apply(X_raw, W_raw, pcs, gt_normals, gt_inst, gt_bb, gt_axes, gt_centers, tuple(weights),
                           norm_eig)


class Backbone(torch.autograd.Function):
    """models/pointnet_extrusion.py:37-66 as one autograd node: forward records a tape, backward walks it with the
    csrc/backward.cu kernels.  `params` are net.parameters() in order, passed only so that autograd routes their
    gradients; BatchNorm running statistics are updated in place by the forward like nn.BatchNorm does."""

    @staticmethod
    def forward(ctx, net, x, fps_start, precision, *params):
        from . import pipeline
        tape = {}
        pipeline.backbone_forward(net, x, fps_start, precision=precision, tape=tape)
        ctx.tape, ctx.net, ctx.precision = tape, net, precision
        ctx.n_params = len(params)
        B, N = tape["B"], tape["N"]
        out = tape["out"]
        return out.reshape(B, N, out.shape[1])

    @staticmethod
    def backward(ctx, d_out):
        from . import backward as bw
        params = list(ctx.net.parameters())
        grads = {id(p): torch.zeros_like(p) for p in params}
        bw.backbone_backward(ctx.tape, d_out, lambda p: grads[id(p)], ctx.precision)
        ctx.tape = None
        return (None, None, None, None) + tuple(grads[id(p)] for p in params)


def backbone_apply(net, x, fps_start=None, precision=None):
    """-> list of (B,N,o_i) views, differentiable w.r.t. net.parameters()."""
    out = Backbone.apply(net, x, fps_start, precision, *list(net.parameters()))
    results, c0 = [], 0
    for fc in net.fc2:
        o = fc.weight.shape[0]
        results.append(out[:, :, c0:c0 + o])
        c0 += o
    return results


class SetAbstractionFn(torch.autograd.Function):
    """One PointNetSetAbstraction level (models/pointnet_util.py:181-207) as an autograd node: differentiable w.r.t.
    the input features (rows) and the level's parameters (not w.r.t. the coordinates, which are data)."""

    @staticmethod
    def forward(ctx, sa, xyz, feats, start, *params):
        from . import pipeline
        tape = {}
        new_xyz, out = pipeline.set_abstraction(sa, xyz, feats, start, tape=tape)
        ctx.tape, ctx.sa = tape, sa
        ctx.mark_non_differentiable(new_xyz)
        return new_xyz, out

    @staticmethod
    def backward(ctx, _d_xyz, d_out):
        from . import backward as bw
        from . import pipeline
        params = list(ctx.sa.parameters())
        grads = {id(p): torch.zeros_like(p) for p in params}
        prec = pipeline._PRECISIONS[pipeline.get_precision()]
        # the backward kernels work in place on the gradient buffer: never on the tensor autograd handed us (it may
        # be shared with a sibling branch or be the user's `backward(gradient=...)` argument)
        d_feats = bw._sa_backward(ctx.tape, d_out.clone(memory_format=torch.contiguous_format),
                                  lambda p: grads[id(p)], prec)
        ctx.tape = None
        return (None, None, d_feats, None) + tuple(grads[id(p)] for p in params)


class SetAbstractionMsgFn(torch.autograd.Function):
    """PointNetSetAbstractionMsg (models/pointnet_util.py:228-267) as an autograd node: one shared FPS, one ball query +
    grouped MLP stack + max-pool per scale, outputs concatenated.  Differentiable w.r.t. the input features (rows)
    and the module's parameters."""

    @staticmethod
    def forward(ctx, sa, xyz, feats, start, *params):
        from . import pipeline
        B, N, _ = xyz.shape
        D = 0 if feats is None else feats.shape[1]
        _, new_xyz = ops.fps(xyz, sa.npoint, start)
        outs, scales = [], []
        for i, radius in enumerate(sa.radius_list):
            idx = ops.ball_query(radius, sa.nsample_list[i], xyz, new_xyz)
            # the Msg variant concatenates [features, centred xyz] (:246-249): gather, then swap the two blocks
            rows = ops.group(xyz, feats, new_xyz, idx, ldo=3 + D)
            if D:
                rows = torch.cat([rows[:, 3:], rows[:, :3]], dim=1).contiguous()
            layers = []
            outs.append(pipeline.mlp_stack(rows, 3 + D, sa.conv_blocks[i], sa.bn_blocks[i], sa.training,
                                           pool_group=sa.nsample_list[i], tag=f"sa_msg{i}", tape=layers))
            scales.append((idx, layers))
        ctx.sa, ctx.scales, ctx.dims = sa, scales, (B, N, D)
        ctx.mark_non_differentiable(new_xyz)
        return new_xyz, torch.cat(outs, dim=1)

    @staticmethod
    def backward(ctx, _d_xyz, d_out):
        from . import backward as bw
        from . import pipeline
        B, N, D = ctx.dims
        params = list(ctx.sa.parameters())
        grads = {id(p): torch.zeros_like(p) for p in params}
        prec = pipeline._PRECISIONS[pipeline.get_precision()]
        d_feats, c0 = None, 0
        for idx, layers in ctx.scales:
            C = layers[-1]["conv"].weight.shape[0]
            d_rows = bw._stack_backward(layers, d_out[:, c0:c0 + C].contiguous(), lambda p: grads[id(p)], prec,
                                        need_input_grad=D > 0)
            c0 += C
            if D:     # rows are [features | xyz]: p2c_group_bwd wants the features behind three xyz columns
                d = ops.group_bwd(torch.cat([d_rows[:, D:D + 3], d_rows[:, :D]], dim=1).contiguous(), idx, B, N, D)
                d_feats = d if d_feats is None else d_feats + d
        ctx.scales = None
        return (None, None, d_feats, None) + tuple(grads[id(p)] for p in params)


class FeaturePropagationFn(torch.autograd.Function):
    """One PointNetFeaturePropagation level (models/pointnet_util.py:283-320): differentiable w.r.t. both feature
    inputs (rows) and the level's parameters."""

    @staticmethod
    def forward(ctx, fp, xyz1, xyz2, feats1, feats2, *params):
        from . import pipeline
        tape = {}
        out = pipeline.feature_propagation(fp, xyz1, xyz2, feats1, feats2, materialize=True, tape=tape)
        ctx.tape, ctx.fp = tape, fp
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import backward as bw
        from . import pipeline
        params = list(ctx.fp.parameters())
        grads = {id(p): torch.zeros_like(p) for p in params}
        prec = pipeline._PRECISIONS[pipeline.get_precision()]
        d_f1, d_f2 = bw._fp_backward(ctx.tape, d_out.clone(memory_format=torch.contiguous_format),   # (see above)
                                     lambda p: grads[id(p)], prec)
        ctx.tape = None
        return (None, None, None, d_f1, d_f2) + tuple(grads[id(p)] for p in params)


class SketchProject(torch.autograd.Function):
    """sketch_implicit_projection (data_utils.py:1014-1146) with the gradient of the projected normals w.r.t. X - the
    with-sketch trainer projects PREDICTED normals (train_Point2Cyl.py:549).  P, labels, axes and centres are data."""

    @staticmethod
    def forward(ctx, X, P, lists, counts, rand_idx, axes, centers, S, zero_tol):
        P_proj, X_proj, scales, found, sel, R = ops.sketch_project(P, X, lists, counts, rand_idx, axes, centers, S,
                                                                   zero_tol, want_sel=True)
        ctx.save_for_backward(sel, R)
        ctx.shape = (P.shape[0], P.shape[1])
        ctx.mark_non_differentiable(P_proj, scales, found)
        return P_proj, X_proj, scales, found

    @staticmethod
    def backward(ctx, _dP, dX_proj, _ds, _df):
        sel, R = ctx.saved_tensors
        B, N = ctx.shape
        return (ops.sketch_project_bwd(dX_proj, sel, R, B, N),) + (None,) * 8
