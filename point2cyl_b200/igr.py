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

  backward         training THROUGH that input gradient (the reference builds it with create_graph=True and calls
                   backward()): two more sweeps, again without autograd (`implicit_backward`; the closed form is
                   restated and checked against torch's double backward in oracle/igr_oracle.py
                   implicit_backward_closed_form).  Going up, the adjoint of the reverse sweep: r_bar_{i+1} =
                   s_i * (r_bar_i W_i^T) and the extra pre-activation gradient beta (1 - s_i) a_i * (r_bar_i W_i^T) leave
                   ONE p2c_linear_act_bwd launch per layer (op 3); going down, the ordinary data gradient with that extra
                   term injected (op 4).  Weight gradients: two p2c_wgrad launches per layer (a_i^T r_bar_i and
                   delta_i^T p_i), bias gradients from the second; the 512 -> 1 output layer through column sums.
                   `ImplicitValueAndGrad` / `Encoder` are the torch.autograd nodes on top (dropin/IGR/network.py).
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
    P: Optional[List[Tensor]] = None        # input rows of every layer (hidden and output), for the weight gradients
    wsplit: Optional[List[Tensor]] = None   # tf32 hi/lo split of the hidden layers' weights
    wsplit_t: Optional[List[Tensor]] = None  # ... of W_i^T restricted to the h-part, i = 1..L-1 (reverse sweep)
    A: Optional[List[Tensor]] = None        # a_i = d f / d z_i of the reverse sweep, i = 0..L-1
    blocks: Optional[List[Tuple[int, int, int]]] = None   # (first row, instances, points per instance) per pts block


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
    Pin: List[Tensor] = []
    h = X0
    for i in range(L):
        Pin.append(h)
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
    Pin.append(h)
    f = torch.empty(R, 1, dtype=torch.float32, device=dev)
    wl = lins[L].weight
    _lib.set_tag("igr.out")
    call("p2c_igr_rowdots", ptr(h), h.stride(0), R, wl.shape[1], ptr(wl), wl.stride(0), 1, 1, ptr(lins[L].bias), 1.0, ptr(f),
         1, 0, stream_ptr())
    if not want_grad:
        return f, None
    blocks = None
    if pts_list is not None:
        blocks, r0 = [], 0
        for p in pts_list:
            blocks.append((r0, I, p.shape[1]))
            r0 += I * p.shape[1]
    return f, ImplicitContext(net, S, R, d_in, P=Pin, wsplit=wsplit, blocks=blocks)


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
    ctx.wsplit_t = wsplit_t
    ctx.A = [None] * L
    ctx.A[L - 1] = a
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
        ctx.A[i - 1] = a
    _lib.set_tag("igr.tail0")
    tail(a, lins[0], 1.0)
    return g


def implicit_backward(ctx: ImplicitContext, f_bar: Optional[Tensor], g_bar: Optional[Tensor], want_dx: bool = True):
    """Backward of (f, g = d f / d x[:, -2:]) of one implicit_forward + implicit_input_gradient pair, without autograd
    (oracle/igr_oracle.py implicit_backward_closed_form; what torch does for IGR/network.py:8-17 with create_graph=True
    followed by backward(), train_Point2Cyl.py:608-672).  f_bar (R, 1) = dL/df, g_bar (R, 2) = dL/dg; either may be None.
    -> ({parameter: gradient} for every lin{i}.weight / .bias, dx (R, d_in) or None)."""
    net = ctx.net
    lins = _layers(net)
    L = len(lins) - 1
    beta = _beta(net)
    skip = set(int(s) for s in net.skip_in)
    dev, R, d_in = lins[0].weight.device, ctx.R, ctx.d_in
    S, P, A = ctx.S, ctx.P, ctx.A
    if P is None or (g_bar is not None and A is None):
        raise _lib.P2CError("implicit_backward: the context carries no tape (forward with want_grad, then the input "
                            "gradient sweep)")
    f32 = dict(dtype=torch.float32, device=dev)
    gW = [torch.zeros_like(l.weight) for l in lins]
    gb = [torch.zeros_like(l.bias) for l in lins]
    in_L = lins[L].weight.shape[1]
    # ---- going up: adjoint of the reverse sweep (only the input-gradient losses feed it) ----
    ZE: List[Optional[Tensor]] = [None] * L
    if g_bar is not None:
        g_bar = g_bar.reshape(R, 2).float()
        RB = torch.zeros(R, ops.pad4(d_in), **f32)
        RB[:, d_in - 2:d_in].copy_(g_bar)
        for i in range(L):
            out_i, in_i = lins[i].weight.shape
            _lib.set_tag(f"igr.bwd.up{i}.wgrad")
            ops.wgrad(A[i][:, :out_i], RB[:, :in_i], in_i, gW[i], None)
            if (i + 1) in skip:
                in_n = lins[i + 1].weight.shape[1]
                nxt = torch.zeros(R, ops.pad4(in_n), **f32)
                nxt[:, in_n - 2:in_n].copy_(g_bar / RT2)
                osc = 1.0 / RT2
            else:
                nxt = torch.empty(R, ops.pad4(out_i), **f32)
                if nxt.shape[1] != out_i:
                    nxt[:, out_i:].zero_()
                osc = 1.0
            ZE[i] = torch.empty(R, ops.pad4(out_i), **f32)
            if ZE[i].shape[1] != out_i:
                ZE[i][:, out_i:].zero_()
            _lib.set_tag(f"igr.bwd.up{i}")
            ops.linear_act_bwd(RB[:, :in_i], ctx.wsplit[i], out_i, in_i, 3, S[i][:, :out_i], A[i][:, :out_i],
                               nxt[:, :out_i], beta=beta, oscale=osc, Z=ZE[i][:, :out_i])
            RB = nxt
        _lib.set_tag("igr.bwd.up.out")
        call("p2c_igr_colsums", ptr(RB), RB.stride(0), R, in_L, None, 0, 1, 1.0, ptr(gW[L]), gW[L].stride(0), stream_ptr())
    # ---- going down: backward of the forward sweep, the extra pre-activation gradients injected ----
    wl = lins[L].weight
    fb = None
    if f_bar is not None:
        fb = f_bar.reshape(R).float().contiguous()
        _lib.set_tag("igr.bwd.out")
        call("p2c_igr_colsums", ptr(P[L]), P[L].stride(0), R, in_L, ptr(fb), 1, 1, 1.0, ptr(gW[L]), gW[L].stride(0),
             stream_ptr())
        gb[L] += fb.sum()
    out_h = lins[L - 1].weight.shape[0]
    D = torch.empty(R, ops.pad4(out_h), **f32)
    if D.shape[1] != out_h:
        D[:, out_h:].zero_()
    _lib.set_tag("igr.bwd.seed")
    if out_h % 4 == 0:
        call("p2c_igr_seed_delta", ptr(ZE[L - 1]), 0 if ZE[L - 1] is None else ZE[L - 1].stride(0), ptr(S[L - 1]),
             S[L - 1].stride(0), ptr(fb), ptr(wl), R, out_h, ptr(D), D.stride(0), stream_ptr())
    else:
        _seed_delta_slow(ZE[L - 1], S[L - 1], fb, wl, D, out_h)
    wsplit_t = ctx.wsplit_t
    if wsplit_t is None:
        srcs = []
        for i in range(1, L):
            in_i = lins[i].weight.shape[1]
            srcs.append(lins[i].weight[:, :in_i - d_in if i in skip else in_i])
        wsplit_t = ops.split_tf32_multi(srcs, transposed=[True] * len(srcs)) if srcs else []
    dx = torch.zeros(R, ops.pad4(d_in), **f32) if want_dx else None
    for i in range(L - 1, -1, -1):
        out_i, in_i = lins[i].weight.shape
        _lib.set_tag(f"igr.bwd.down{i}.wgrad")
        ops.wgrad(D[:, :out_i], P[i][:, :in_i], in_i, gW[i], gb[i])
        n_h = in_i - d_in if i in skip else in_i
        if want_dx and (i == 0 or i in skip):
            # the columns of p_bar_i that are the network input itself: x enters layer 0 and, scaled, every skip layer
            src = lins[i].weight if i == 0 else lins[i].weight[:, n_h:]
            wt = ops.split_tf32_multi([src], transposed=[True])[0]
            part = torch.empty(R, ops.pad4(d_in), **f32)
            _lib.set_tag(f"igr.bwd.down{i}.dx")
            ops.linear_act(D[:, :out_i], wt, None, d_in, out_i, op=0, out=part[:, :d_in])
            dx[:, :d_in].add_(part[:, :d_in], alpha=1.0 if i == 0 else 1.0 / RT2)
        if i == 0:
            break
        osc = 1.0 / RT2 if i in skip else 1.0
        _lib.set_tag(f"igr.bwd.down{i}")
        if ZE[i - 1] is not None:
            nxt = ZE[i - 1]                               # in place: Y = acc * s * osc + ZE
            ops.linear_act_bwd(D[:, :out_i], wsplit_t[i - 1], n_h, out_i, 4, S[i - 1][:, :n_h], nxt[:, :n_h],
                               nxt[:, :n_h], beta=beta, oscale=osc)
        else:
            nxt = torch.empty(R, S[i - 1].shape[1], **f32)
            if nxt.shape[1] != n_h:
                nxt[:, n_h:].zero_()
            ops.linear_act(D[:, :out_i], wsplit_t[i - 1], None, n_h, out_i, op=2, oscale=osc, out=nxt[:, :n_h],
                           mul=S[i - 1][:, :n_h])
        D = nxt
    grads = {}
    for i, l in enumerate(lins):
        grads[l.weight] = gW[i]
        grads[l.bias] = gb[i]
    return grads, (dx[:, :d_in] if want_dx else None)


def _seed_delta_slow(ZE, S, fb, wl, D, C):
    """Hidden width not a multiple of 4 (never the case for the trainer's 512-wide network): plain torch."""
    D[:, :C].zero_()
    if ZE is not None:
        D[:, :C].add_(ZE[:, :C])
    if fb is not None:
        D[:, :C].add_(S[:, :C] * fb[:, None] * wl.reshape(1, -1)[:, :C])


def latent_grad(ctx: ImplicitContext, dx: Tensor, want_pts: bool = False):
    """Backward of add_latent over the pts blocks of a forward made from (latent, pts): dx (R, d_in) -> d latent (I, E)
    [, list of d pts (I, S_j, 2)]."""
    if ctx.blocks is None:
        raise _lib.P2CError("latent_grad: the forward was not made from (latent, pts)")
    E = ctx.d_in - 2
    I = ctx.blocks[0][1]
    dlat = torch.zeros(I, E, dtype=torch.float32, device=dx.device)
    dpts = []
    for r0, inst, Sn in ctx.blocks:
        blk = dx[r0:r0 + inst * Sn]
        dp = torch.empty(inst, Sn, 2, dtype=torch.float32, device=dx.device) if want_pts else None
        call("p2c_igr_latent_grad", ptr(blk), blk.stride(0), inst, Sn, E, 1.0, ptr(dlat), ptr(dp), 0, stream_ptr())
        dpts.append(dp)
    return (dlat, dpts) if want_pts else dlat


def _net_params(net) -> List[Tensor]:
    return [p for l in _layers(net) for p in (l.weight, l.bias)]


def _backward_outputs(ictx: ImplicitContext, f_bar, g_bar, need_x: bool, params):
    grads, dx = implicit_backward(ictx, f_bar, g_bar, want_dx=need_x)
    return dx, tuple(grads[p] if p.requires_grad else None for p in params)


class _ImplicitValueFn(torch.autograd.Function):
    """f = ImplicitNet(x) as an autograd node (IGR/network.py:67-92); the context goes out through `holder`."""

    @staticmethod
    def forward(ctx, net, holder, x, *params):
        f, ictx = implicit_forward(net, x=x, want_grad=True)
        holder["ctx"] = ictx
        ctx.ictx, ctx.params = ictx, params
        return f

    @staticmethod
    def backward(ctx, f_bar):
        dx, gp = _backward_outputs(ctx.ictx, f_bar.contiguous(), None, ctx.needs_input_grad[2], ctx.params)
        return (None, None, dx) + gp


class _ImplicitGradFn(torch.autograd.Function):
    """g = gradient(x, f) (IGR/network.py:8-17, create_graph=True) as an autograd node: its backward is the adjoint of
    the reverse sweep followed by the backward of the forward sweep (implicit_backward with f_bar = None)."""

    @staticmethod
    def forward(ctx, ictx, x, *params):
        ctx.ictx, ctx.params = ictx, params
        return implicit_input_gradient(ictx)

    @staticmethod
    def backward(ctx, g_bar):
        dx, gp = _backward_outputs(ctx.ictx, None, g_bar.contiguous(), ctx.needs_input_grad[1], ctx.params)
        return (None, dx) + gp


def implicit_net_forward(net, x: Tensor) -> Tensor:
    """What the drop-in ImplicitNet.forward calls: values always; with grad mode on and anything requiring grad the
    result is an autograd node whose backward runs the closed-form sweeps."""
    params = _net_params(net)
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        holder: dict = {}
        f = _ImplicitValueFn.apply(net, holder, x, *params)
        f._p2c_igr = holder["ctx"]
        return f
    f, ctx = implicit_forward(net, x=x, want_grad=bool(x.requires_grad))
    if ctx is not None:
        f._p2c_igr = ctx
    return f


def gradient(inputs: Tensor, outputs: Tensor) -> Tensor:
    """IGR/network.py:8-17 for outputs produced by the kernels' ImplicitNet forward: the closed-form reverse sweep; with
    grad mode on it is itself differentiable (the reference's create_graph=True)."""
    ctx = getattr(outputs, "_p2c_igr", None)
    if ctx is None:
        raise _lib.P2CError("gradient(): `outputs` does not come from a point2cyl_b200 ImplicitNet forward with "
                            "`inputs.requires_grad` set (the closed-form sweep needs the saved softplus' values)")
    params = _net_params(ctx.net)
    if torch.is_grad_enabled() and outputs.requires_grad:
        return _ImplicitGradFn.apply(ctx, inputs, *params)
    return implicit_input_gradient(ctx)


class _AddLatentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, latent_codes):
        ctx.shape = tuple(points.shape)
        ctx.E = latent_codes.shape[1]
        return _add_latent_values(points, latent_codes)

    @staticmethod
    def backward(ctx, dx):
        I, S, _ = ctx.shape
        dx = dx.contiguous()
        dlat = torch.zeros(I, ctx.E, dtype=torch.float32, device=dx.device)
        dp = torch.empty(I, S, 2, dtype=torch.float32, device=dx.device) if ctx.needs_input_grad[0] else None
        call("p2c_igr_latent_grad", ptr(dx), dx.stride(0), I, S, ctx.E, 1.0, ptr(dlat), ptr(dp), 0, stream_ptr())
        return dp, dlat


def add_latent(points: Tensor, latent_codes: Tensor) -> Tensor:
    """IGR/network.py:200-206: (I,S,d=2), (I,E) -> (I*S, E+2), latent first; differentiable (the latent codes come from
    the trained PointNetEncoder)."""
    if points.shape[2] == 2 and torch.is_grad_enabled() and (points.requires_grad or latent_codes.requires_grad):
        return _AddLatentFn.apply(points, latent_codes)
    return _add_latent_values(points, latent_codes)


def _add_latent_values(points: Tensor, latent_codes: Tensor) -> Tensor:
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


def _encoder_modules(enc):
    convs = [enc.mlp1[0], enc.mlp1[3], enc.mlp2[0], enc.mlp2[3], enc.mlp2[6]]
    bns = [enc.mlp1[1], enc.mlp1[4], enc.mlp2[1], enc.mlp2[4], enc.mlp2[7]]
    return convs, bns


def _encoder_stack(enc, x: Tensor, tape: Optional[list]):
    """x (I, S, >= C) -> per-group maxima (I * S / pg, 1024) of the five Conv1d+BN+ReLU layers, pg."""
    _lib.need_cuda(x)
    I, S, _ = x.shape
    C = enc.input_channels
    rows = torch.zeros(I * S, ops.pad4(C), dtype=torch.float32, device=x.device)
    rows[:, :C].copy_(x[:, :, :C].reshape(I * S, C))
    convs, bns = _encoder_modules(enc)
    pg = next((g for g in (128, 64, 32) if S % g == 0), 0)
    if pg == 0:
        raise _lib.P2CError(f"PointNetEncoder: {S} points per instance is not a multiple of 32")
    return pipeline.mlp_stack(rows, C, convs, bns, enc.training, pool_group=pg, tag="igr.enc", tape=tape), pg


class _EncoderStackFn(torch.autograd.Function):
    """The per-point layers of PointNetEncoder as one autograd node (backward: point2cyl_b200.backward's layer walk)."""

    @staticmethod
    def forward(ctx, enc, x, *params):
        tape: list = []
        pooled, pg = _encoder_stack(enc, x, tape)
        ctx.tape, ctx.enc, ctx.params, ctx.xshape = tape, enc, params, tuple(x.shape)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        from . import backward as bw
        grads = {id(p): torch.zeros_like(p) for p in ctx.params}
        prec = pipeline._PRECISIONS[pipeline.get_precision()]
        need_x = ctx.needs_input_grad[1]
        d_rows = bw._stack_backward(ctx.tape, d_pooled.clone(memory_format=torch.contiguous_format),
                                    lambda p: grads[id(p)], prec, need_input_grad=need_x)
        ctx.tape = None
        dx = None
        if need_x:
            I, S, Cx = ctx.xshape
            C = ctx.enc.input_channels
            dx = torch.zeros(I, S, Cx, dtype=torch.float32, device=d_pooled.device)
            dx[:, :, :C].copy_(d_rows[:, :C].reshape(I, S, C))
        return (None, dx) + tuple(grads[id(p)] for p in ctx.params)


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the fp32 kernels, differentiable (the encoder's fc on <= a few hundred rows)."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return ops.linear(x.contiguous(), W, b, precision=_lib.PREC_FP32)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        dW = torch.zeros_like(W)
        db = torch.zeros(W.shape[0], dtype=torch.float32, device=W.device)
        ops.wgrad(dy, x.contiguous(), W.shape[1], dW, db, precision=_lib.PREC_FP32)
        dx = ops.linear(dy, W.t().contiguous(), None, precision=_lib.PREC_FP32) if ctx.needs_input_grad[0] else None
        return dx, dW, db


def encoder_forward(enc, x: Tensor) -> Tensor:
    """PointNetEncoder.forward (IGR/network.py:162-174): x (I, S, >= C) -> unit latent codes (I, E).  Conv1d+BN+ReLU x5
    on the per-point MLP kernels (BatchNorm statistics and the max over points in the layer epilogues), Linear,
    F.normalize.  With grad mode on the result is differentiable w.r.t. the encoder's parameters and x."""
    I, S, _ = x.shape
    convs, bns = _encoder_modules(enc)
    params = [p for m in convs + bns for p in (m.weight, m.bias)]
    taped = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    if taped:
        pooled = _EncoderStackFn.apply(enc, x, *params)            # (I*S/pg, 1024)
    else:
        pooled, _ = _encoder_stack(enc, x, None)
    groups = pooled.shape[0] // I
    if groups != 1:
        pooled = pooled.reshape(I, groups, pooled.shape[1]).max(dim=1)[0]   # max of the per-group maxima
    _lib.set_tag("igr.enc.fc")
    if taped:
        h = _LinearFn.apply(pooled, enc.fc.weight, enc.fc.bias)
    else:
        h = ops.linear(pooled.contiguous(), enc.fc.weight, enc.fc.bias, precision=_lib.PREC_FP32)
    return F.normalize(h)


# ---- the loss block ------------------------------------------------------------------------------------------------


def masked_instance_mean(loss: Tensor, mask_gt: Tensor) -> Tensor:
    """losses.py:83-88 (reduce_mean_masked_instance) on (B, K)."""
    kept = torch.where(mask_gt, loss, torch.zeros_like(loss)).sum(dim=1)
    cnt = mask_gt.sum(dim=1).to(loss.dtype)
    return torch.where(cnt > 0, kept / cnt, torch.zeros_like(kept))


def _sketch_terms(net, latent, sk_pnts, sk_normals, off_pnts):
    """-> per-instance loss terms (I, 3) = {mean |f|, SALD normal term, eikonal term}, f_on, g_on, g_off, context."""
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
    return terms, f_on, g_on, g_off, ctx


class _SketchTermsFn(torch.autograd.Function):
    """The per-instance loss terms of train_Point2Cyl.py:608-647 as ONE autograd node over the whole implicit block
    (forward sweep, input-gradient sweep, term reduction); backward = p2c_igr_loss_terms_bwd + implicit_backward +
    the backward of add_latent.  Differentiable w.r.t. the latent codes and ImplicitNet's parameters."""

    @staticmethod
    def forward(ctx, net, latent, sk_pnts, sk_normals, off_pnts, *params):
        terms, f_on, g_on, g_off, ictx = _sketch_terms(net, latent, sk_pnts, sk_normals, off_pnts)
        ctx.ictx, ctx.params = ictx, params
        ctx.save_for_backward(f_on, g_on, g_off, sk_normals)
        ctx.mark_non_differentiable(f_on, g_on, g_off)
        return terms, f_on, g_on, g_off

    @staticmethod
    def backward(ctx, dterms, *_):
        f_on, g_on, g_off, nrm = ctx.saved_tensors
        ictx = ctx.ictx
        I, S = nrm.shape[0], nrm.shape[1]
        So = g_off.shape[0] // I
        R = ictx.R
        f_bar = torch.zeros(R, 1, dtype=torch.float32, device=f_on.device)     # off-surface values carry no loss
        g_bar = torch.empty(R, 2, dtype=torch.float32, device=f_on.device)
        _lib.set_tag("igr.loss.bwd")
        call("p2c_igr_loss_terms_bwd", ptr(f_on), ptr(g_on), ptr(nrm.contiguous().float()), ptr(g_off), I, S, So,
             ptr(dterms.contiguous().float()), ptr(f_bar), ptr(g_bar), ptr(g_bar[I * S:]), stream_ptr())
        need_lat = ctx.needs_input_grad[1]
        grads, dx = implicit_backward(ictx, f_bar, g_bar, want_dx=need_lat)
        dlat = latent_grad(ictx, dx) if need_lat else None
        ctx.ictx = None
        return (None, dlat, None, None, None) + tuple(grads[p] if p.requires_grad else None for p in ctx.params)


def sketch_loss_block(net, latent: Tensor, latent_gt: Tensor, sk_pnts: Tensor, sk_normals: Tensor, off_pnts: Tensor,
                      mask_gt: Tensor, is_l2: bool = False) -> Dict[str, Tensor]:
    """train_Point2Cyl.py:608-672: manifold + 0.1 eikonal + SALD-normal + latent loss of B*K sketch instances.
    latent / latent_gt (B*K, E); sk_pnts, sk_normals (B*K, S, 2); off_pnts (B*K, S_off, 2) from the sampler; mask_gt
    (B, K).  On- and off-surface points go through the network in ONE sweep.  With grad mode on, im_loss.backward()
    fills the gradients of ImplicitNet's parameters and of `latent` (closed-form sweeps, no autograd through the net)."""
    B, K = mask_gt.shape
    params = _net_params(net)
    if torch.is_grad_enabled() and (latent.requires_grad or any(p.requires_grad for p in params)):
        terms, f_on, g_on, g_off = _SketchTermsFn.apply(net, latent, sk_pnts, sk_normals, off_pnts, *params)
    else:
        terms, f_on, g_on, g_off, _ = _sketch_terms(net, latent, sk_pnts, sk_normals, off_pnts)
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
