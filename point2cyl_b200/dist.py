"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo on CPU).

The forward+loss path needs NO data-path collective: clouds are independent (FPS, ball query, per-cloud loss
statistics, per-sample assignment), so a global batch is split into contiguous per-rank shards.
What a training step adds is a single sum all-reduce of one flat fp32 gradient buffer (1.4 M elements for heads
[3,16]); train-mode BatchNorm statistics stay per replica exactly like the reference, which has no SyncBN.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str = None) -> Tuple[int, int, int]:
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `total` clouds owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    B = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(B, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def reduce_max(x: float, device=None) -> float:
    """max over ranks of a host scalar (device-timed durations are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


class FlatGradAllReduce:
    """One all-reduce per step over a single flat buffer that the parameters' .grad tensors are views of."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def sync(self) -> None:
        """Average gradients over ranks (sum all-reduce, then 1/world), in place."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())


def broadcast_buffers(module: torch.nn.Module, src: int = 0) -> None:
    """Replicas keep their own BatchNorm running statistics (like the reference); before a checkpoint is
    written they are made identical to rank `src`'s."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for b in module.buffers():
            dist.broadcast(b, src)
