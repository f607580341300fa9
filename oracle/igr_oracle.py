"""CPU oracle for the implicit sketch network of the with-sketch trainer.  TEST INFRASTRUCTURE ONLY.

Groundwork for the next widening step: there is NO CUDA path for this block yet; this file
pins the arithmetic the kernels will have to reproduce.  Like p2c_oracle.py it is a plain torch-CPU restatement,
imported only by tests/ (and tests/golden/make_golden.py), never by the product package.

Parity status: PINNED by tests/golden/igr_*.npz, produced by the reference's own IGR/network.py + IGR/sampler.py and
the loss lines of train_Point2Cyl.py run through oracle/ref_shim.load_igr() (tests/golden/make_golden.py), checked
in tests/test_oracle_igr.py.

Reference (paths relative to the upstream root):
  IGR/network.py:8-17      gradient(): d out / d in by autograd with create_graph, last two columns
  IGR/network.py:20-92     ImplicitNet: Linear stack, softplus(beta=100), skip concatenation / sqrt(2)
  IGR/network.py:132-174   PointNetEncoder: Conv1d+BN+ReLU x5, max over points, Linear, F.normalize
  IGR/network.py:200-206   add_latent
  IGR/sampler.py:19-37     NormalPerPoint.get_points
  train_Point2Cyl.py:598-672   latent codes, manifold / eikonal / SALD-normal / latent losses
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

IMPLICIT_DIMS = (512,) * 8          # train_Point2Cyl.py:264
IMPLICIT_SKIP = (4,)
LATENT_SIZE = 256                   # :260
D_IN = 2                            # :259
ENCODER_WIDTHS = (64, 64, 64, 128, 1024)   # IGR/network.py:141-159 (mlp1: two layers, mlp2: three)


# ---- ImplicitNet ------------------------------------------------------------------------------------------------


def implicit_layer_shapes(d_in: int = D_IN + LATENT_SIZE, dims: Sequence[int] = IMPLICIT_DIMS,
                          skip_in: Sequence[int] = IMPLICIT_SKIP) -> List[Tuple[int, int]]:
    """(fan_in, fan_out) per Linear, IGR/network.py:32-45: the layer BEFORE a skip layer is narrower by d_in so that
    the concatenation restores the nominal width."""
    full = [d_in] + list(dims) + [1]
    return [(full[i], full[i + 1] - d_in if (i + 1) in skip_in else full[i + 1]) for i in range(len(full) - 1)]


def implicit_init(d_in: int = D_IN + LATENT_SIZE, dims: Sequence[int] = IMPLICIT_DIMS,
                  skip_in: Sequence[int] = IMPLICIT_SKIP, radius_init: float = 1.0, seed: int = 0) -> Dict[str, Tensor]:
    """Geometric initialisation (IGR/network.py:48-58): hidden weights N(0, sqrt(2/out)), zero bias; last layer
    N(sqrt(pi/in), 1e-5), bias -radius.  Seeded with a private generator (the VALUES are not the reference's random
    stream; tests load the same dict into the reference module)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    shapes = implicit_layer_shapes(d_in, dims, skip_in)
    for i, (fi, fo) in enumerate(shapes):
        if i == len(shapes) - 1:
            sd[f"lin{i}.weight"] = math.sqrt(math.pi) / math.sqrt(fi) + 1e-5 * torch.randn(fo, fi, generator=g)
            sd[f"lin{i}.bias"] = torch.full((fo,), -float(radius_init))
        else:
            sd[f"lin{i}.weight"] = math.sqrt(2.0) / math.sqrt(fo) * torch.randn(fo, fi, generator=g)
            sd[f"lin{i}.bias"] = torch.zeros(fo)
    return sd


def softplus(x: Tensor, beta: float = 100.0) -> Tensor:
    """nn.Softplus(beta): log(1 + exp(beta x)) / beta, and exactly x where beta x > 20 (torch's threshold)."""
    return F.softplus(x, beta=beta, threshold=20.0)


def implicit_forward(sd: Dict[str, Tensor], x: Tensor, skip_in: Sequence[int] = IMPLICIT_SKIP,
                     beta: float = 100.0) -> Tensor:
    """IGR/network.py:67-92: (R, d_in) -> (R, 1).  No activation after the last layer."""
    n_layers = len([k for k in sd if k.endswith(".weight")])
    h = x
    for i in range(n_layers):
        if i in skip_in:
            h = torch.cat([h, x], dim=-1) / math.sqrt(2.0)
        h = F.linear(h, sd[f"lin{i}.weight"], sd[f"lin{i}.bias"])
        if i < n_layers - 1:
            h = softplus(h, beta)
    return h


def gradient(inputs: Tensor, outputs: Tensor) -> Tensor:
    """IGR/network.py:8-17: d(sum outputs)/d inputs, last TWO columns (the 2-D point; the latent columns are dropped),
    differentiable again (create_graph) because the eikonal / normal losses are trained through it."""
    g = torch.autograd.grad(outputs=outputs, inputs=inputs, grad_outputs=torch.ones_like(outputs), create_graph=True,
                            retain_graph=True, only_inputs=True)[0]
    return g[:, -2:]


def implicit_forward_with_input_grad(sd: Dict[str, Tensor], x: Tensor, skip_in: Sequence[int] = IMPLICIT_SKIP,
                                     beta: float = 100.0) -> Tuple[Tensor, Tensor]:
    """The closed form a kernel computes instead of autograd: forward keeping the pre-activations z_i, then
    d f / d x = J_0^T ... with softplus'(z) = sigmoid(beta z) (1 where beta z > 20), the skip layer contributing
    1/sqrt(2) of its trailing d_in columns directly.  -> (f (R,1), df/dx[:, -2:] (R,2)).  No autograd."""
    n_layers = len([k for k in sd if k.endswith(".weight")])
    d_in = x.shape[1]
    h, zs = x, []
    for i in range(n_layers):
        if i in skip_in:
            h = torch.cat([h, x], dim=-1) / math.sqrt(2.0)
        z = F.linear(h, sd[f"lin{i}.weight"], sd[f"lin{i}.bias"])
        zs.append(z)
        h = softplus(z, beta) if i < n_layers - 1 else z
    g = torch.ones_like(zs[-1])                                   # d f / d z_last
    gx = torch.zeros_like(x)
    for i in range(n_layers - 1, -1, -1):
        if i < n_layers - 1:
            bz = beta * zs[i]
            g = g * torch.where(bz > 20.0, torch.ones_like(bz), torch.sigmoid(bz))
        g = g @ sd[f"lin{i}.weight"]                              # d f / d (input of layer i)
        if i in skip_in:
            g = g / math.sqrt(2.0)
            gx = gx + g[:, -d_in:]
            g = g[:, :-d_in]
    gx = gx + g
    return zs[-1], gx[:, -2:]


def implicit_backward_closed_form(sd: Dict[str, Tensor], x: Tensor, f_bar: Tensor, g_bar: Tensor,
                                  skip_in: Sequence[int] = IMPLICIT_SKIP, beta: float = 100.0):
    """Backward of (f, g = df/dx[:, -2:]) WITHOUT autograd: the four sweeps a kernel implementation runs.
    Given the upstream gradients f_bar (R,1) = dL/df and g_bar (R,2) = dL/dg, returns ({name: grad} for every weight
    and bias, dL/dx (R, d_in)).  Notation (row vectors): p_i input of layer i (h_i, or cat(h_i, x)/sqrt 2 at a skip
    layer), z_i = p_i W_i^T + b_i, h_{i+1} = softplus(z_i), s_i = softplus'(z_i) = sigmoid(beta z_i),
    t_i = softplus''(z_i) = beta s_i (1 - s_i) (s = 1, t = 0 where beta z > 20, torch's threshold).
      1. forward            keep p_i, z_i
      2. reverse (for g)    a_L = 1, r_i = a_i W_i, q_i = d f / d h_i (r_i, or its leading part / sqrt 2 at a skip),
                            a_{i-1} = s_{i-1} * q_i,  g_x = r_0 + sum_skip r_s[:, n_h:] / sqrt 2
      3. adjoint of 2.      r_bar_0 = g_bar (+ g_bar / sqrt 2 into the trailing columns of r_bar_s); going up:
                            a_bar_i = r_bar_i W_i^T, dW_i += a_i^T r_bar_i, q_bar_{i+1} = a_bar_i * s_i,
                            z_extra_i = a_bar_i * q_{i+1} * t_i
      4. backward of 1.     delta_L = f_bar, dW_i += delta_i^T p_i, db_i += sum delta_i, p_bar_i = delta_i W_i,
                            delta_{i-1} = z_extra_{i-1} + s_{i-1} * h_bar_i,  dx collects p_bar_0 and the skip parts.
    This is what torch.autograd does for `gradient(..., create_graph=True)` followed by backward(); checked against it
    in tests/test_oracle_igr.py."""
    L = len([k for k in sd if k.endswith(".weight")]) - 1          # index of the last layer
    d_in = x.shape[1]
    rt2 = math.sqrt(2.0)
    W = [sd[f"lin{i}.weight"] for i in range(L + 1)]
    # 1. forward
    p, z, h = [], [], x
    for i in range(L + 1):
        pi = torch.cat([h, x], dim=-1) / rt2 if i in skip_in else h
        zi = pi @ W[i].t() + sd[f"lin{i}.bias"]
        p.append(pi)
        z.append(zi)
        h = softplus(zi, beta) if i < L else zi
    s, t = [], []
    for i in range(L):
        bz = beta * z[i]
        sg = torch.sigmoid(bz)
        lin = bz > 20.0
        s.append(torch.where(lin, torch.ones_like(sg), sg))
        t.append(torch.where(lin, torch.zeros_like(sg), beta * sg * (1.0 - sg)))
    # 2. reverse sweep
    a = [None] * (L + 1)
    q = [None] * (L + 1)                                          # q[i] = d f / d h_i, i >= 1
    a[L] = torch.ones_like(z[L])
    for i in range(L, 0, -1):
        r = a[i] @ W[i]
        q[i] = r[:, :r.shape[1] - d_in] / rt2 if i in skip_in else r
        a[i - 1] = s[i - 1] * q[i]
    # 3. adjoint of the reverse sweep
    gx_bar = torch.zeros_like(x)
    gx_bar[:, -2:] = g_bar
    grads = {f"lin{i}.weight": torch.zeros_like(W[i]) for i in range(L + 1)}
    z_extra = [None] * L
    r_bar = gx_bar
    for i in range(L + 1):
        if i in skip_in and i > 0:                                # trailing columns of r_i feed g_x directly
            r_bar = torch.cat([r_bar, gx_bar / rt2], dim=-1)
        grads[f"lin{i}.weight"] += a[i].t() @ r_bar
        if i == L:
            break
        a_bar = r_bar @ W[i].t()
        z_extra[i] = a_bar * q[i + 1] * t[i]
        q_bar = a_bar * s[i]
        r_bar = q_bar / rt2 if (i + 1) in skip_in else q_bar      # leading (h) part of r_bar_{i+1}
    # 4. backward of the forward, with the extra pre-activation gradients injected
    dx = torch.zeros_like(x)
    delta = f_bar
    for i in range(L, -1, -1):
        grads[f"lin{i}.weight"] += delta.t() @ p[i]
        grads[f"lin{i}.bias"] = delta.sum(dim=0)
        p_bar = delta @ W[i]
        if i in skip_in:
            dx += p_bar[:, -d_in:] / rt2
            h_bar = p_bar[:, :-d_in] / rt2
        else:
            h_bar = p_bar
        if i == 0:
            dx += h_bar
        else:
            delta = z_extra[i - 1] + s[i - 1] * h_bar
    return grads, dx


# ---- PointNetEncoder ---------------------------------------------------------------------------------------------


def encoder_init(embedding_size: int = LATENT_SIZE, input_channels: int = 2 * D_IN, seed: int = 0) -> Dict[str, Tensor]:
    """state_dict of PointNetEncoder(embedding_size, D_IN, with_normals=True) (IGR/network.py:132-160): keys
    mlp1.{0,3}, mlp2.{0,3,6} convs with BatchNorm at +1, fc."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = input_channels
    slots = [("mlp1", 0), ("mlp1", 3), ("mlp2", 0), ("mlp2", 3), ("mlp2", 6)]
    for (blk, j), cout in zip(slots, ENCODER_WIDTHS):
        bound = 1.0 / math.sqrt(cin)
        sd[f"{blk}.{j}.weight"] = (torch.rand(cout, cin, 1, generator=g) * 2 - 1) * bound
        sd[f"{blk}.{j}.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        sd[f"{blk}.{j + 1}.weight"] = 1.0 + 0.2 * (torch.rand(cout, generator=g) - 0.5)
        sd[f"{blk}.{j + 1}.bias"] = 0.2 * (torch.rand(cout, generator=g) - 0.5)
        sd[f"{blk}.{j + 1}.running_mean"] = 0.1 * (torch.rand(cout, generator=g) - 0.5)
        sd[f"{blk}.{j + 1}.running_var"] = 0.5 + torch.rand(cout, generator=g)
        sd[f"{blk}.{j + 1}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        cin = cout
    bound = 1.0 / math.sqrt(cin)
    sd["fc.weight"] = (torch.rand(embedding_size, cin, generator=g) * 2 - 1) * bound
    sd["fc.bias"] = (torch.rand(embedding_size, generator=g) * 2 - 1) * bound
    return sd


def encoder_forward(sd: Dict[str, Tensor], x: Tensor, input_channels: int = 2 * D_IN, training: bool = True,
                    eps: float = 1e-5) -> Tensor:
    """IGR/network.py:162-174: x (R, S, >=C) -> unit latent codes (R, E).  Train-mode BatchNorm uses the batch
    statistics over (R, S) (running buffers are not updated here: the oracle is functional)."""
    h = x[:, :, :input_channels].transpose(2, 1)
    for blk, j in (("mlp1", 0), ("mlp1", 3), ("mlp2", 0), ("mlp2", 3), ("mlp2", 6)):
        h = F.conv1d(h, sd[f"{blk}.{j}.weight"], sd[f"{blk}.{j}.bias"])
        if training:
            h = F.batch_norm(h, None, None, sd[f"{blk}.{j + 1}.weight"], sd[f"{blk}.{j + 1}.bias"], True, 0.0, eps)
        else:
            h = F.batch_norm(h, sd[f"{blk}.{j + 1}.running_mean"], sd[f"{blk}.{j + 1}.running_var"],
                             sd[f"{blk}.{j + 1}.weight"], sd[f"{blk}.{j + 1}.bias"], False, 0.0, eps)
        h = F.relu(h)
    h = h.max(dim=2)[0]                                            # F.max_pool1d(x, num_points)
    h = F.linear(h, sd["fc.weight"], sd["fc.bias"])
    return F.normalize(h)


# ---- sampling and the loss block -----------------------------------------------------------------------------------


def add_latent(points: Tensor, latent_codes: Tensor) -> Tensor:
    """IGR/network.py:200-206: (R,S,d), (R,E) -> (R*S, E+d), latent first."""
    R, S, d = points.shape
    lat = latent_codes[:, None, :].expand(R, S, latent_codes.shape[1]).reshape(R * S, -1)
    return torch.cat([lat, points.reshape(R * S, d)], dim=1)


def sample_off_surface(pc: Tensor, global_sigma: float = 1.8, local_sigma: float = 0.01) -> Tensor:
    """NormalPerPoint.get_points (IGR/sampler.py:25-37): every point jittered by N(0, local_sigma) plus S//8 uniform
    points in [-global_sigma, global_sigma]^d; consumes the global generator in the reference's order (randn, rand)."""
    R, S, d = pc.shape
    local = pc + torch.randn_like(pc) * local_sigma
    glob = torch.rand(R, S // 8, d) * (global_sigma * 2) - global_sigma
    return torch.cat([local, glob], dim=1)


def masked_instance_mean(loss: Tensor, mask_gt: Tensor) -> Tensor:
    """losses.py:83-88 (reduce_mean_masked_instance): mean over existing instances per cloud, 0 when there are none."""
    kept = torch.where(mask_gt, loss, torch.zeros_like(loss)).sum(dim=1)
    cnt = mask_gt.sum(dim=1).to(loss.dtype)
    return torch.where(cnt > 0, kept / cnt, torch.zeros_like(kept))


def sketch_loss_block(net_sd: Dict[str, Tensor], latent: Tensor, latent_gt: Tensor, sk_pnts: Tensor, sk_normals: Tensor,
                      off_pnts: Tensor, mask_gt: Tensor, is_l2: bool = False, analytic: bool = False) -> Dict[str, Tensor]:
    """train_Point2Cyl.py:608-672.  latent / latent_gt (B*K, E); sk_pnts, sk_normals (B*K, S, 2); off_pnts
    (B*K, S + S//8, 2) from sample_off_surface; mask_gt (B, K) bool.
    -> im_loss (= manifold + 0.1 eikonal + normals + latent) and its terms, plus the per-point predictions."""
    B, K = mask_gt.shape
    x_on = add_latent(sk_pnts, latent)
    x_off = add_latent(off_pnts, latent)
    if analytic:
        f_on, g_on = implicit_forward_with_input_grad(net_sd, x_on)
        _, g_off = implicit_forward_with_input_grad(net_sd, x_off)
    else:
        if not x_on.requires_grad:
            x_on.requires_grad_()
        if not x_off.requires_grad:
            x_off.requires_grad_()
        f_on = implicit_forward(net_sd, x_on)
        f_off = implicit_forward(net_sd, x_off)
        g_on, g_off = gradient(x_on, f_on), gradient(x_off, f_off)
    pred = f_on.reshape(B, K, -1, 1)
    g_off4, g_on4 = g_off.reshape(B, K, -1, 2), g_on.reshape(B, K, -1, 2)
    nrm = sk_normals.reshape(B, K, -1, 2)
    mnfld = masked_instance_mean(pred.abs().mean(dim=-1).mean(dim=-1), mask_gt).mean()
    eik = masked_instance_mean(((g_off4.norm(2, dim=-1) - 1) ** 2).mean(dim=-1), mask_gt).mean()
    sald = torch.minimum((g_on4 - nrm).norm(2, dim=-1), (g_on4 + nrm).norm(2, dim=-1)).mean(dim=-1)
    sald = masked_instance_mean(sald, mask_gt).mean()
    lat, lat_gt = latent.reshape(B, K, -1), latent_gt.reshape(B, K, -1)
    if is_l2:
        latent_loss = masked_instance_mean(torch.square(lat - lat_gt).sum(dim=-1), mask_gt).mean()
    else:
        latent_loss = masked_instance_mean(1.0 - (lat * lat_gt).sum(dim=-1), mask_gt).mean()
    im = mnfld + 0.1 * eik + 1.0 * sald + latent_loss
    return dict(im_loss=im, mnfld_loss=mnfld, grad_loss=eik, normals_loss=sald, latent_loss=latent_loss,
                sk_pred=f_on, mnfld_grad=g_on, nonmnfld_grad=g_off)
