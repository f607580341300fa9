"""point2cyl_b200 — B200-native (sm_100a) forward+loss hot path of Point2Cyl.

Public surface:
  point2cyl_b200.dropin.*        drop-in modules with the reference's names and signatures (differentiable)
  point2cyl_b200.pipeline        backbone_forward / loss_forward / forward_loss on device tensors
  point2cyl_b200.forward_loss_host   the same call from HOST buffers (H2D in, loss D2H out)
  point2cyl_b200.graph           GraphedForwardLoss: CUDA-graph replay of the step, host copies overlapped
  point2cyl_b200.autograd        torch.autograd.Function wrappers whose backward runs the backward kernels
  point2cyl_b200.train           Trainer / GraphedTrainer: forward + loss + backward + all-reduce + Adam, flat buffers
  point2cyl_b200.dist            one-process-per-GPU helpers (shards, reductions, buffer broadcast)
  point2cyl_b200.ops             one wrapper per C-ABI entry point (include/point2cyl.h)
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

__all__ = ["forward_loss_host", "pin_batch"]

BATCH_KEYS = ("pcs", "normals", "inst", "bb", "axes", "centers")


def pin_batch(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Page-lock a host batch (what a DataLoader(pin_memory=True) would hand over)."""
    return {k: batch[k].contiguous().pin_memory() for k in BATCH_KEYS}


def forward_loss_host(net, host_batch: Dict[str, torch.Tensor], device="cuda", fps_start=None,
                      weights=(1.0,) * 5, norm_eig: bool = False,
                      precision: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Forward + loss of one batch whose tensors live in (pinned) host memory.

    Mirrors the reference loop body train_Point2Cyl_without_sketch.py:226-353: host->device copies of
    the six batch tensors, backbone forward, loss block; returns the six loss scalars on the HOST
    (`losses`: total, normal, miou, bb, axis, center) plus the device-side outputs.
    """
    from . import pipeline
    dev_batch = {k: host_batch[k].to(device, non_blocking=True) for k in BATCH_KEYS}
    out = pipeline.forward_loss(net, dev_batch, fps_start, weights, norm_eig, precision)
    out["losses_host"] = out["losses"].cpu()
    return out


def h2d_bytes(host_batch: Dict[str, torch.Tensor]) -> int:
    return sum(host_batch[k].numel() * host_batch[k].element_size() for k in BATCH_KEYS)
