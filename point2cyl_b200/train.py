"""One data-parallel training step of Point2Cyl on libp2c.so kernels:

    forward (tape) -> loss kernels -> loss backward kernels -> backbone backward kernels
    -> ONE all-reduce of the flat gradient buffer (NCCL over NVLink on the GPU box) -> fused Adam kernel

This is the loop body of train_Point2Cyl_without_sketch.py:244-369 without torch autograd in the way: parameters and
their gradients are views of two flat fp32 buffers, so the collective and the optimiser are one launch each.
The drop-in modules remain usable with plain `loss.backward()` + torch.optim (point2cyl_b200.autograd); this class
is the fast path the benchmark times.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.distributed as dist

from . import backward as bw
from . import ops, pipeline

Tensor = torch.Tensor


class Trainer:
    def __init__(self, net: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, weights: Sequence[float] = (1.0,) * 5, norm_eig: bool = False,
                 precision: Optional[str] = None):
        self.net = net
        self.params = [p for p in net.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        # parameters and gradients become views of flat buffers (state_dict keys / shapes are unchanged)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.weights, self.norm_eig, self.precision = tuple(weights), norm_eig, precision
        self.step_count = 0
        self._ones = None

    def world(self) -> int:
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    @torch.no_grad()
    def forward_backward(self, batch: Dict[str, Tensor], fps_start=None) -> Dict[str, Tensor]:
        """Gradients of the total loss into the flat buffer (zeroed first).  Returns the loss dict."""
        self.flat_grad.zero_()
        tape: Dict = {}
        X_raw, W_raw = pipeline.backbone_forward(self.net, batch["pcs"], fps_start, precision=self.precision, tape=tape)
        B, N, twoK = W_raw.shape
        K = twoK // 2
        stats = ops.segfit_stats(X_raw, W_raw, batch["pcs"], batch["normals"], batch["inst"], batch["bb"], K)
        cost, n_gt = ops.segfit_cost(stats, K)
        match = ops.hungarian(cost, n_gt)
        bb_sum = ops.bb_loss_sums(W_raw, batch["bb"], match, n_gt, K)
        losses, E_AX, centers, per_seg, per_cloud = ops.loss_finalize(
            stats, bb_sum, match, n_gt, batch["axes"], batch["centers"], N, K, self.norm_eig, self.weights)
        if self._ones is None or self._ones.device != stats.device:
            self._ones = torch.tensor(self.weights, dtype=torch.float32, device=stats.device)
        dstats = ops.loss_backward_coef(stats, match, n_gt, batch["axes"], batch["centers"], self._ones, N, K,
                                        self.norm_eig)
        C_out = 3 + 2 * K                                 # rows padded to 16 bytes: TMA operand of the heads' wgrad
        d_buf = torch.empty(B * N, ops.pad4(C_out), dtype=torch.float32, device=stats.device)
        d_out = ops.segfit_backward(X_raw, W_raw, batch["pcs"], batch["normals"], batch["inst"], batch["bb"], dstats,
                                    match, n_gt, self._ones, K, out=d_buf[:, :C_out])
        bw.backbone_backward(tape, d_out, lambda p: p.grad, self.precision)
        return dict(total=losses[0], losses=losses, matching_indices=match, E_AX=E_AX, centers=centers)

    @torch.no_grad()
    def step(self, batch: Dict[str, Tensor], fps_start=None) -> Dict[str, Tensor]:
        """forward + loss + backward + gradient all-reduce + Adam.  Gradients are averaged over ranks like DDP."""
        out = self.forward_backward(batch, fps_start)
        world = self.world()
        if world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        self.step_count += 1
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.step_count,
                      self.betas, self.eps, self.weight_decay, grad_scale=1.0 / world)
        return out
