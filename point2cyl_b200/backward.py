"""Backward pass of the backbone over the tape that pipeline.backbone_forward(tape=...) records.

Mirror of the forward's stage list, walked in reverse (what torch autograd does for
models/pointnet_extrusion.py:37-66 in the reference):

    heads -> fc1 -> fp1 -> fp2 -> fp3 -> sa3 -> sa2 -> sa1

Every per-point operation is a libp2c.so kernel (csrc/backward.cu + the forward's tcgen05 p2c_linear for the data
gradients); torch is used for allocation, the tiny weight transposes and the accumulation of the two gradient
contributions that the skip connections create (l1, l2).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import _lib, ops

Tensor = torch.Tensor
GradOf = Callable[[torch.nn.Parameter], Tensor]


def _wprec(precision: int) -> int:
    """wgrad follows the forward's precision: tensor cores (3xTF32) unless the fp32 SIMT path was asked for."""
    return _lib.PREC_FP32 if precision == _lib.PREC_FP32 else _lib.PREC_3XTF32


def _w2(conv) -> Tensor:
    return conv.weight.reshape(conv.weight.shape[0], -1)


def _dgrad(dY: Tensor, W: Tensor, precision: int) -> Tensor:
    """dA_prev (M, K) = dY (M, N) @ W (N, K): the forward kernel on the transposed weight."""
    Wt = W.t().contiguous()                      # (K, N): 'output channels' = K, reduction over N
    Kout, Nred = Wt.shape
    wsplit = ops.weight_operand(dY, Wt, Kout, Nred, False, 0, precision)
    return ops.linear(dY, Wt, None, K=Nred, precision=precision, w_split=wsplit)


def _layer_backward(rec: dict, dA: Tensor, grad_of: GradOf, precision: int, need_input_grad: bool):
    """One [conv -> BN -> ReLU] layer.  dA: gradient of the layer's post-BN/ReLU output — (M, C), or (M/pool, C) for
    a max-pooled layer.  Returns (dY raw-output gradient, dA_prev or None)."""
    bn, conv = rec["bn"], rec["conv"]
    scale, shift = rec["scale"], rec["shift"]
    C_ = rec["Y"].shape[1]
    fused = ops.bn_bwd_fusable(C_)       # the per-channel coefficients are evaluated inside the applying kernel
    if rec["pool"]:
        sums = ops.pool_bwd_reduce(dA, rec["Ymax"], rec["Ymin"], scale, shift)
    else:
        sums = ops.bn_bwd_reduce(dA, rec["Y"], scale, shift)
    coef = None
    if not fused:
        coef = ops.bn_bwd_coef(sums, rec["M"], bn.weight, rec["mean"], rec["invstd"], rec["batch_stats"],
                               grad_of(bn.weight), grad_of(bn.bias))
    if rec["pool"]:
        if fused:
            dY = ops.pool_bwd_apply_fused(dA, rec["Ymax"], rec["Ymin"], rec["Y"], scale, shift, sums, rec["M"],
                                          bn.weight, rec["mean"], rec["invstd"], rec["batch_stats"],
                                          grad_of(bn.weight), grad_of(bn.bias), rec["pool"])
        else:
            dY = ops.pool_bwd_apply(dA, rec["Ymax"], rec["Ymin"], rec["Y"], scale, shift, coef, rec["pool"])
    else:
        out = dA if dA.is_contiguous() else None
        if fused:
            dY = ops.bn_bwd_apply_fused(dA, rec["Y"], scale, shift, sums, rec["M"], bn.weight, rec["mean"],
                                        rec["invstd"], rec["batch_stats"], grad_of(bn.weight), grad_of(bn.bias), out=out)
        else:
            dY = ops.bn_bwd_apply(dA, rec["Y"], scale, shift, coef, out=out)
    if rec["fused_first"]:
        return dY, None
    aff = rec["in_aff"]
    ops.wgrad(dY, rec["X"], rec["K"], grad_of(conv.weight), grad_of(conv.bias),
              None if aff is None else aff.scale, None if aff is None else aff.shift, precision=_wprec(precision))
    dA_prev = _dgrad(dY, _w2(conv), precision) if need_input_grad else None
    return dY, dA_prev


def _stack_backward(layers, dA: Tensor, grad_of: GradOf, precision: int, need_input_grad: bool):
    """Reverse walk over an mlp_stack's records.  Returns the gradient w.r.t. the stack input rows (None when not
    needed) or, when the first layer is the fused gather+conv, that layer's raw-output gradient dY0."""
    for i in range(len(layers) - 1, -1, -1):
        rec = layers[i]
        dY, dA = _layer_backward(rec, dA, grad_of, precision, need_input_grad or i > 0)
        if rec["fused_first"]:
            return dY
    return dA


def _sa_backward(rec: dict, d_out: Tensor, grad_of: GradOf, precision: int) -> Optional[Tensor]:
    """Set-abstraction level: d_out (B*S, C_last) -> gradient w.r.t. the level's input features (B*N, D) or None."""
    layers = rec["layers"]
    D = rec["D"]
    if not rec["fused"]:
        d_rows = _stack_backward(layers, d_out, grad_of, precision, need_input_grad=D > 0)
        if D == 0:
            return None
        if rec.get("grouped"):
            # un-fused ball-query level (first width other than 64 / 128): scatter the rows back to their source points
            return ops.group_bwd(d_rows, rec["gidx"], rec["B"], rec["N"], D)
        # group-all level: rows = [xyz | feats] in point order; the feature gradient is a column slice
        return d_rows[:, 3:3 + D]
    dY0 = _stack_backward(layers, d_out, grad_of, precision, need_input_grad=False)
    conv0 = layers[0]["conv"]
    gW0 = grad_of(conv0.weight).reshape(conv0.weight.shape[0], -1)
    feats = rec["feats"]
    dQf = None
    if feats is not None:
        dQf = torch.zeros(rec["B"] * rec["N"], dY0.shape[1], dtype=torch.float32, device=dY0.device)
    ops.sa_first_bwd(dY0, rec["xyz"], rec["new_xyz"], rec["gidx"], dQf, gW0, grad_of(conv0.bias))
    if feats is None:
        return None
    # Qf = feats @ W0[:, 3:].T  (computed once per source point in the forward)
    ops.wgrad(dQf, feats, D, gW0[:, 3:], None, precision=_wprec(precision))
    return _dgrad(dQf, _w2(conv0)[:, 3:], precision)


def _fp_backward(rec: dict, d_out: Tensor, grad_of: GradOf, precision: int):
    """Feature-propagation level: d_out (B*N, C_last) -> (d feats1 (B*N, D1) view or None, d feats2 (B*S, D2))."""
    d_buf = _stack_backward(rec["layers"], d_out, grad_of, precision, need_input_grad=True)
    D1 = rec["D1"]
    d_f1 = d_buf[:, :D1] if D1 else None
    d_f2 = ops.three_nn_interp_bwd(d_buf[:, D1:], rec["nn_idx"], rec["nn_w"], rec["B"], rec["N"], rec["S"])
    return d_f1, d_f2


def backbone_backward(tape: Dict, d_out: Tensor, grad_of: GradOf, precision: Optional[str] = None) -> None:
    """d_out: (B*N, sum(output_sizes)) gradient of the heads' output rows.  Parameter gradients are ACCUMULATED into
    the tensors grad_of(param) returns (same shape as the parameter)."""
    from . import pipeline
    prec = pipeline._PRECISIONS[precision or pipeline.get_precision()]
    B, N = tape["B"], tape["N"]
    head = tape["head"]
    if d_out.dim() == 3:
        d_out = d_out.contiguous().reshape(B * N, -1)
    d_rows = d_out                                   # (B*N, C) rows, possibly with a padded row stride
    _lib.set_tag("bwd.head")
    # heads: p2c_head_bwd gives the data gradient and re-materialises the heads' input relu(bn1(fc1)) * dropout mask;
    # their weights / biases then come from p2c_wgrad on plain row matrices (tensor cores when d_out rows are 16-byte
    # aligned - the Trainer allocates them padded; a contiguous autograd gradient is copied into a padded buffer)
    Wcat = head["Wcat"]
    aff_h = head["aff_h"]
    dA, A_h = ops.head_bwd(d_rows, head["mask_cf"], Wcat, B, N, head["h"], aff_h.scale, aff_h.shift,
                           seed=head.get("seed"))
    if d_rows.stride(0) % 4:
        padded = torch.zeros(B * N, ops.pad4(d_rows.shape[1]), dtype=torch.float32, device=d_rows.device)
        padded[:, :d_rows.shape[1]].copy_(d_rows)
        d_rows = padded[:, :d_rows.shape[1]]
    gWcat = torch.zeros_like(Wcat)
    gbcat = torch.zeros(Wcat.shape[0], dtype=torch.float32, device=Wcat.device)
    ops.wgrad(d_rows, A_h, Wcat.shape[1], gWcat, gbcat, precision=_wprec(prec))
    del A_h
    c0 = 0
    for fc in head["fc2"]:
        o = fc.weight.shape[0]
        grad_of(fc.weight).reshape(o, -1).add_(gWcat[c0:c0 + o])
        grad_of(fc.bias).add_(gbcat[c0:c0 + o])
        c0 += o
    dA = _stack_backward(head["layers"], dA, grad_of, prec, need_input_grad=True)        # fc1 / bn1
    _lib.set_tag("bwd.fp1")
    d_feats0, d_l5 = _fp_backward(tape["fp1"], dA, grad_of, prec)                        # d_feats0: input normals, unused
    _lib.set_tag("bwd.fp2")
    d_l1_skip, d_l4 = _fp_backward(tape["fp2"], d_l5, grad_of, prec)
    _lib.set_tag("bwd.fp3")
    d_l2_skip, d_l3 = _fp_backward(tape["fp3"], d_l4, grad_of, prec)
    _lib.set_tag("bwd.sa3")
    d_l2 = _sa_backward(tape["sa3"], d_l3, grad_of, prec)
    d_l2 = (d_l2 + d_l2_skip).contiguous()
    _lib.set_tag("bwd.sa2")
    d_l1 = _sa_backward(tape["sa2"], d_l2, grad_of, prec)
    d_l1 = (d_l1 + d_l1_skip).contiguous()
    _lib.set_tag("bwd.sa1")
    _sa_backward(tape["sa1"], d_l1, grad_of, prec)
