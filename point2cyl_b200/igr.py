"""The implicit sketch network of the with-sketch trainer on the libp2c.so kernels.

Reference: IGR/network.py:8-17 (`gradient`), :20-92 (`ImplicitNet`), :132-174 (`PointNetEncoder`), :200-206
(`add_latent`), IGR/sampler.py:19-37 (`NormalPerPoint`) and the loss lines train_Point2Cyl.py:598-672.

`ImplicitNet` is eight 512-wide Linear+Softplus(100) layers over ~10^5..10^6 rows: unlike the backbone this block is
TENSOR bound (3.7 MFLOP per row forward, the same again for the input gradient).  It runs as

  forward sweep    one p2c_linear_act launch per hidden layer (tcgen05, 3xTF32, weights pre-split by ONE
                   p2c_split_tf32_multi launch): Z = P W^T + b in tensor memory, the epilogue writes H = softplus(Z)
                   (the next layer's input; the skip layer's share of the concatenation lands pre-scaled by 1/sqrt 2 in
                   the concat buffer) and S = sigmoid(100 Z) = softplus'(Z);
  reverse sweep    d f / d x WITHOUT autograd (the closed form of oracle/igr_oracle.py, checked there against
                   torch.autograd.grad(create_graph=True)): a_{L-1} = S_{L-1} * W_L, then one p2c_linear_act launch per
                   layer on W_i^T with the multiplier epilogue a_{i-1} = S_{i-1} * (a_i W_i); only the two columns of
                   the 2-D point are kept (network.py:17), so the first layer and the skip layer contribute through
                   two thin row-dot kernels instead of full GEMMs.

Forward values only: the second-order backward (training the implicit network through its own input gradient) is the
four-sweep closed form of oracle/igr_oracle.implicit_backward_closed_form and is not built yet.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib, ops, pipeline
from ._lib import call, ptr, stream_ptr

Tensor = torch.Tensor
RT2 = math.sqrt(2.0)
_debug_layers = None          # tests/tools/igr_debug.py: list that receives every hidden layer's output buffer


def _layers(net) -> List[torch.nn.Linear]:
    return [getattr(net, f"lin{i}") for i in range(net.num_layers - 1)]


@dataclass
class ImplicitContext:
    """What the reverse sweep needs from a forward sweep (kept on the output tensor for `gradient`)."""
    net: object
    S: List[Tensor]            # sigmoid(beta z_i) per hidden layer, (R, pad4(out_i))
    R: int
    d_in: int


def _beta(net) -> float:
    act = getattr(net, "activation", None)
    if not isinstance(act, torch.nn.Softplus):
        raise _lib.P2CError("ImplicitNet on the kernels needs the Softplus activation (beta > 0), as the trainer uses it")
    if float(act.threshold) != 20.0:
        raise _lib.P2CError("Softplus threshold other than torch's default 20 is not supported")
    return float(act.beta)


def implicit_forward(net, x: Optional[Tensor] = None, latent: Optional[Tensor] = None, pts: Optional[Tensor] = None,
                     want_grad: bool = True):
    """ImplicitNet.forward (IGR/network.py:67-92).  Either x (R, d_in), or latent (I, E) + pts (I, S, 2) [a list of such
    pts tensors is concatenated row-wise: on-surface then off-surface points share one sweep] which add_latent
    (:200-206) would turn into x.  -> (f (R, 1), ImplicitContext or None)."""
    lins = _layers(net)
    L = len(lins) - 1                                   # index of the output layer
    beta = _beta(net)
    skip = set(int(s) for s in net.skip_in)
    dev = lins[0].weight.device
    d_in = lins[0].weight.shape[1]
    if L in skip or 0 in skip:
        raise _lib.P2CError("a skip connection into the first or the output layer is not supported")
    # ---- layer inputs: P[i] (R, pad4(in_i)); the input of a skip layer is [h | x] / sqrt(2) ----
    if x is not None:
        _lib.need_cuda(x)
        R = x.shape[0]
        X0 = torch.zeros(R, ops.pad4(d_in), dtype=torch.float32, device=dev)
        X0[:, :d_in].copy_(x)
        pts_list = None
    else:
        pts_list = list(pts) if isinstance(pts, (list, tuple)) else [pts]
        I, E = latent.shape
        if E + 2 != d_in:
            raise _lib.P2CError(f"latent size {E} + 2 != d_in {d_in}")
        R = sum(I * p.shape[1] for p in pts_list)
        X0 = torch.empty(R, ops.pad4(d_in), dtype=torch.float32, device=dev)
    P: Dict[int, Tensor] = {0: X0}
    for s in skip:
        P[s] = torch.zeros(R, ops.pad4(lins[s].weight.shape[1]), dtype=torch.float32, device=dev)
    if x is not None:
        for s in skip:
            n_h = lins[s].weight.shape[1] - d_in
            P[s][:, n_h:n_h + d_in].copy_(X0[:, :d_in] / RT2)
    else:
        lat = latent.contiguous().float()
        r0 = 0
        skips = sorted(skip)
        for p in pts_list:
            Sn = p.shape[1]
            rows = I * Sn
            p2 = p.contiguous().float().reshape(rows, 2)
            first = skips[0] if skips else None
            Pt = P[first][r0:r0 + rows] if first is not None else None
            call("p2c_igr_add_latent", ptr(lat), ptr(p2), rows, Sn, E, ptr(X0[r0:r0 + rows]), X0.stride(0), ptr(Pt),
                 0 if Pt is None else Pt.stride(0), 0 if Pt is None else lins[first].weight.shape[1] - d_in,
                 1.0 / RT2, stream_ptr())
            r0 += rows
        for s in skips[1:]:
            n_h = lins[s].weight.shape[1] - d_in
            P[s][:, n_h:n_h + d_in].copy_(P[skips[0]][:, lins[skips[0]].weight.shape[1] - d_in:][:, :d_in])
    # ---- every hidden layer's weights split into tf32 hi / lo by one launch ----
    wsplit = ops.split_tf32_multi([l.weight for l in lins[:L]])
    S: List[Tensor] = []
    h = X0
    for i in range(L):
        out_i, in_i = lins[i].weight.shape
        if (i + 1) in skip:
            Y, osc = P[i + 1], 1.0 / RT2                # leading columns of the concat buffer, pre-scaled
        else:
            Y, osc = torch.empty(R, ops.pad4(out_i), dtype=torch.float32, device=dev), 1.0
            if Y.shape[1] != out_i:
                Y[:, out_i:].zero_()
        Si = None
        if want_grad:
            Si = torch.empty(R, ops.pad4(out_i), dtype=torch.float32, device=dev)
            S.append(Si)
        _lib.set_tag(f"igr.fwd{i}")
        ops.linear_act(h, wsplit[i], lins[i].bias, out_i, in_i, op=1, beta=beta, oscale=osc, out=Y[:, :out_i],
                       S=None if Si is None else Si[:, :out_i])
        h = Y
        if _debug_layers is not None:
            _debug_layers.append(Y)
    f = torch.empty(R, 1, dtype=torch.float32, device=dev)
    wl = lins[L].weight
    _lib.set_tag("igr.out")
    call("p2c_igr_rowdots", ptr(h), h.stride(0), R, wl.shape[1], ptr(wl), wl.stride(0), 1, 1, ptr(lins[L].bias), 1.0, ptr(f),
         1, 0, stream_ptr())
    return f, (ImplicitContext(net, S, R, d_in) if want_grad else None)


def implicit_input_gradient(ctx: ImplicitContext) -> Tensor:
    """d f / d x restricted to its last two columns - what `gradient(inputs, outputs)` returns (IGR/network.py:8-17) -
    by the reverse sweep over the saved softplus'(z_i)."""
    net = ctx.net
    lins = _layers(net)
    L = len(lins) - 1
    skip = set(int(s) for s in net.skip_in)
    dev, R, d_in = lins[0].weight.device, ctx.R, ctx.d_in
    S = ctx.S
    wl = lins[L].weight                                   # (1, in_L)
    a = torch.empty(R, S[L - 1].shape[1], dtype=torch.float32, device=dev)
    _lib.set_tag("igr.seed")
    call("p2c_igr_scale_cols", ptr(S[L - 1]), S[L - 1].stride(0), ptr(wl), R, wl.shape[1], ptr(a), a.stride(0),
         stream_ptr())
    # W_i^T restricted to the h-part of layer i's input, split by one launch
    srcs = []
    for i in range(1, L):
        out_i, in_i = lins[i].weight.shape
        n_h = in_i - d_in if i in skip else in_i
        srcs.append(lins[i].weight[:, :n_h])
    wsplit_t = ops.split_tf32_multi(srcs, transposed=[True] * len(srcs)) if srcs else []
    g = torch.empty(R, 2, dtype=torch.float32, device=dev)
    first = True

    def tail(a_i, lin, scale):
        """g (+)= scale * (a_i W_i)[:, last two columns]"""
        nonlocal first
        W = lin.weight
        col = W.shape[1] - 2
        call("p2c_igr_rowdots", ptr(a_i), a_i.stride(0), R, W.shape[0], W.data_ptr() + 4 * col, 1, W.stride(0), 2, None,
             float(scale), ptr(g), 2, 0 if first else 1, stream_ptr())
        first = False

    for i in range(L - 1, 0, -1):
        out_i, in_i = lins[i].weight.shape
        n_h = in_i - d_in if i in skip else in_i
        if i in skip:
            _lib.set_tag(f"igr.tail{i}")
            tail(a, lins[i], 1.0 / RT2)
        nxt = torch.empty(R, S[i - 1].shape[1], dtype=torch.float32, device=dev)
        if nxt.shape[1] != n_h:
            nxt[:, n_h:].zero_()
        _lib.set_tag(f"igr.rev{i}")
        ops.linear_act(a[:, :out_i], wsplit_t[i - 1], None, n_h, out_i, op=2, oscale=(1.0 / RT2 if i in skip else 1.0),
                       out=nxt[:, :n_h], mul=S[i - 1][:, :n_h])
        a = nxt
    _lib.set_tag("igr.tail0")
    tail(a, lins[0], 1.0)
    return g


def gradient(inputs: Tensor, outputs: Tensor) -> Tensor:
    """IGR/network.py:8-17 for outputs produced by the kernels' ImplicitNet forward (values only, see module doc)."""
    ctx = getattr(outputs, "_p2c_igr", None)
    if ctx is None:
        raise _lib.P2CError("gradient(): `outputs` does not come from a point2cyl_b200 ImplicitNet forward with "
                            "`inputs.requires_grad` set (the closed-form sweep needs the saved softplus' values)")
    return implicit_input_gradient(ctx)


def add_latent(points: Tensor, latent_codes: Tensor) -> Tensor:
    """IGR/network.py:200-206: (I,S,d=2), (I,E) -> (I*S, E+2), latent first."""
    I, S, d = points.shape
    E = latent_codes.shape[1]
    if d != 2:
        return torch.cat([latent_codes.unsqueeze(1).repeat(1, S, 1).reshape(I * S, -1), points.reshape(I * S, d)], 1)
    out = torch.empty(I * S, E + 2, dtype=torch.float32, device=points.device)
    if (E + 2) % 4:
        X0 = torch.empty(I * S, ops.pad4(E + 2), dtype=torch.float32, device=points.device)
    else:
        X0 = out
    call("p2c_igr_add_latent", ptr(latent_codes.contiguous().float()), ptr(points.contiguous().float().reshape(I * S, 2)),
         I * S, S, E, ptr(X0), X0.stride(0), None, 0, 0, 1.0, stream_ptr())
    if X0 is not out:
        out.copy_(X0[:, :E + 2])
    return out


# ---- PointNetEncoder ---------------------------------------------------------------------------------------------


def encoder_forward(enc, x: Tensor) -> Tensor:
    """PointNetEncoder.forward (IGR/network.py:162-174): x (I, S, >= C) -> unit latent codes (I, E).  Conv1d+BN+ReLU x5
    on the per-point MLP kernels (BatchNorm statistics and the max over points in the layer epilogues), Linear,
    F.normalize."""
    _lib.need_cuda(x)
    I, S, _ = x.shape
    C = enc.input_channels
    rows = torch.zeros(I * S, ops.pad4(C), dtype=torch.float32, device=x.device)
    rows[:, :C].copy_(x[:, :, :C].reshape(I * S, C))
    convs = [enc.mlp1[0], enc.mlp1[3], enc.mlp2[0], enc.mlp2[3], enc.mlp2[6]]
    bns = [enc.mlp1[1], enc.mlp1[4], enc.mlp2[1], enc.mlp2[4], enc.mlp2[7]]
    pg = next((g for g in (128, 64, 32) if S % g == 0), 0)
    if pg == 0:
        raise _lib.P2CError(f"PointNetEncoder: {S} points per instance is not a multiple of 32")
    pooled = pipeline.mlp_stack(rows, C, convs, bns, enc.training, pool_group=pg, tag="igr.enc")   # (I*S/pg, 1024)
    if S != pg:
        pooled = pooled.reshape(I, S // pg, pooled.shape[1]).amax(dim=1)     # max of the per-128-point maxima
    _lib.set_tag("igr.enc.fc")
    h = ops.linear(pooled.contiguous(), enc.fc.weight, enc.fc.bias, precision=_lib.PREC_FP32)
    return F.normalize(h)


# ---- the loss block ------------------------------------------------------------------------------------------------


def masked_instance_mean(loss: Tensor, mask_gt: Tensor) -> Tensor:
    """losses.py:83-88 (reduce_mean_masked_instance) on (B, K)."""
    kept = torch.where(mask_gt, loss, torch.zeros_like(loss)).sum(dim=1)
    cnt = mask_gt.sum(dim=1).to(loss.dtype)
    return torch.where(cnt > 0, kept / cnt, torch.zeros_like(kept))


def sketch_loss_block(net, latent: Tensor, latent_gt: Tensor, sk_pnts: Tensor, sk_normals: Tensor, off_pnts: Tensor,
                      mask_gt: Tensor, is_l2: bool = False) -> Dict[str, Tensor]:
    """train_Point2Cyl.py:608-672: manifold + 0.1 eikonal + SALD-normal + latent loss of B*K sketch instances.
    latent / latent_gt (B*K, E); sk_pnts, sk_normals (B*K, S, 2); off_pnts (B*K, S_off, 2) from the sampler; mask_gt
    (B, K).  On- and off-surface points go through the network in ONE sweep."""
    B, K = mask_gt.shape
    I, S, _ = sk_pnts.shape
    So = off_pnts.shape[1]
    f, ctx = implicit_forward(net, latent=latent, pts=[sk_pnts, off_pnts], want_grad=True)
    g = implicit_input_gradient(ctx)
    n_on = I * S
    f_on, g_on, g_off = f[:n_on], g[:n_on], g[n_on:]
    terms = torch.empty(I, 3, dtype=torch.float32, device=f.device)
    _lib.set_tag("igr.loss")
    call("p2c_igr_loss_terms", ptr(f_on), ptr(g_on), ptr(sk_normals.contiguous().float()), ptr(g_off), I, S, So,
         ptr(terms), stream_ptr())
    t = terms.reshape(B, K, 3)
    mnfld = masked_instance_mean(t[:, :, 0], mask_gt).mean()
    sald = masked_instance_mean(t[:, :, 1], mask_gt).mean()
    eik = masked_instance_mean(t[:, :, 2], mask_gt).mean()
    lat, lat_gt = latent.reshape(B, K, -1), latent_gt.reshape(B, K, -1)
    if is_l2:
        latent_loss = masked_instance_mean(torch.square(lat - lat_gt).sum(dim=-1), mask_gt).mean()
    else:
        latent_loss = masked_instance_mean(1.0 - (lat * lat_gt).sum(dim=-1), mask_gt).mean()
    im = mnfld + 0.1 * eik + 1.0 * sald + latent_loss
    return dict(im_loss=im, mnfld_loss=mnfld, grad_loss=eik, normals_loss=sald, latent_loss=latent_loss, sk_pred=f_on,
                mnfld_grad=g_on, nonmnfld_grad=g_off)
