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
from . import _lib, ops, pipeline

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
    def _forward(self, pcs: Tensor, fps_start=None, geo=None):
        """Backbone forward with a tape (needs only the point coordinates of the batch; `geo`: its coordinate-only
        stage computed ahead of time by pipeline.geometry_forward)."""
        self.flat_grad.zero_()
        tape: Dict = {}
        X_raw, W_raw = pipeline.backbone_forward(self.net, pcs, fps_start, precision=self.precision, tape=tape, geo=geo)
        return tape, X_raw, W_raw

    @torch.no_grad()
    def _loss_backward(self, batch: Dict[str, Tensor], tape: Dict, X_raw: Tensor, W_raw: Tensor) -> Dict[str, Tensor]:
        """Loss block, its backward and the backbone backward: gradients land in the flat buffer."""
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
    def forward_backward(self, batch: Dict[str, Tensor], fps_start=None) -> Dict[str, Tensor]:
        """Gradients of the total loss into the flat buffer (zeroed first).  Returns the loss dict."""
        tape, X_raw, W_raw = self._forward(batch["pcs"], fps_start)
        return self._loss_backward(batch, tape, X_raw, W_raw)

    @torch.no_grad()
    def step(self, batch: Optional[Dict[str, Tensor]], fps_start=None) -> Dict[str, Tensor]:
        """forward + loss + backward + gradient all-reduce + Adam.  Gradients are averaged over ranks like DDP."""
        out = self.forward_backward(batch, fps_start)
        world = self.world()
        if world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        self.step_count += 1
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.step_count,
                      self.betas, self.eps, self.weight_decay, grad_scale=1.0 / world)
        return out


class GraphedTrainer(Trainer):
    """Trainer whose forward + loss + backward (about 230 launches of short kernels) is captured ONCE into a CUDA
    graph and replayed; the gradient all-reduce and the Adam kernel (its bias correction is a launch argument that
    changes every step) stay outside the graph.  Per-step host work mirrors the reference: the first FPS centroid of
    each level is drawn from the CPU generator (models/pointnet_util.py:75) and copied to the device.

    Re-capture (build a new object) when the batch shape, train/eval mode or the BatchNorm momentum changes."""

    def __init__(self, net, example: Dict[str, Tensor], warmup: int = 2, auto_rebuild: bool = True, **kw):
        super().__init__(net, **kw)
        from . import BATCH_KEYS
        from .graph import StartRing
        dev = self.flat_param.device
        self.keys = BATCH_KEYS
        self.static = {k: torch.empty_like(example[k], device=dev).copy_(example[k]) for k in BATCH_KEYS}
        B, N, _ = self.static["pcs"].shape
        self.B, self.N, self.S1 = B, N, net.sa1.npoint
        self.starts = StartRing(B, (N, self.S1), dev)
        self.start_dev = self.starts.dev
        self.auto_rebuild = auto_rebuild
        # two graphs, like graph.GraphedForwardLoss: the backbone forward needs only the coordinates, so the labels /
        # normals of a host batch are copied on a side stream while it runs
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = torch.cuda.Event()
        self.captures = 0
        self._capture(warmup)

    def _capture(self, warmup: int):
        net, dev = self.net, self.flat_param.device
        self._key = self._state_key()
        # eager warm-up outside the capture (one-time cudaFuncSetAttribute calls, allocator growth); BatchNorm buffers
        # and gradients are restored afterwards so that building the graph does not count as training steps
        buffers = {k: v.clone() for k, v in net.named_buffers()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._draw_starts()
                Trainer.forward_backward(self, self.static, self.start_dev)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._tape = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            tape, X_raw, W_raw = Trainer._forward(self, self.static["pcs"], self.start_dev)
        self.bwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.bwd_graph, pool=self.graph.pool()):
            self.out = Trainer._loss_backward(self, self.static, tape, X_raw, W_raw)
        self._tape = tape                                  # keeps the forward's activations alive for the replay
        with torch.no_grad():
            for k, v in net.named_buffers():
                v.copy_(buffers[k])
        self.captures += 1

    def _state_key(self):
        from .graph import bn_state_key
        return bn_state_key(self.net)

    def rebuild_if_stale(self) -> bool:
        """Re-capture when mode / BatchNorm momentum changed since the last capture (update_momentum in the training
        scripts, train_Point2Cyl_without_sketch.py:357-366).  Returns True if it did."""
        if self._key == self._state_key():
            return False
        torch.cuda.synchronize(self.flat_param.device)
        self._capture(warmup=1)
        return True

    def _draw_starts(self):
        self.starts.draw()

    @torch.no_grad()
    def forward_backward(self, batch: Optional[Dict[str, Tensor]] = None, fps_start=None) -> Dict[str, Tensor]:
        if self._key != self._state_key():
            if not self.auto_rebuild:
                raise RuntimeError("GraphedTrainer: mode or BatchNorm momentum changed since capture; call "
                                   "rebuild_if_stale()")
            self.rebuild_if_stale()
        cur = torch.cuda.current_stream(self.flat_param.device)
        copying = batch is not None and batch["pcs"] is not self.static["pcs"]
        if copying:
            for k in self.keys:                                  # copy_ would silently broadcast e.g. a short last batch
                if tuple(batch[k].shape) != tuple(self.static[k].shape):
                    raise _lib.P2CError(f"GraphedTrainer was captured for {k} of shape {tuple(self.static[k].shape)}, "
                                        f"got {tuple(batch[k].shape)}: build a new one (or use Trainer) for this batch")
            self.static["pcs"].copy_(batch["pcs"], non_blocking=True)
            self.copy_stream.wait_stream(cur)
            with torch.cuda.stream(self.copy_stream):
                for k in self.keys:
                    if k != "pcs":
                        self.static[k].copy_(batch[k], non_blocking=True)
                self.copied.record(self.copy_stream)
        if fps_start is None:
            self._draw_starts()
        else:
            for d, s in zip(self.start_dev, fps_start):
                d.copy_(s, non_blocking=True)
        self.graph.replay()
        if copying:
            cur.wait_event(self.copied)
        self.bwd_graph.replay()
        return self.out


class PipelinedTrainer(Trainer):
    """Training step over a STREAM of batches as a two-stage software pipeline (one batch deep), the training-side
    counterpart of graph.PipelinedForwardLoss:

        coordinate stage of batch i+1 (second stream)  |  layers + loss + backward + all-reduce + Adam of batch i (main stream)

    The coordinate-only stage (pipeline.geometry_forward: FPS and ball query of both levels, the 3-NN searches; ~0.6 ms
    of short, mostly serial kernels on few SMs) depends on the coordinates only - not on the weights the optimiser is
    about to change - so it runs one batch ahead.  Two buffer slots (ping-pong); per slot one CUDA graph for the
    coordinate stage, one for the feature forward (captured with `p2c_set_sm_budget(SMs - geometry_sms)`: the
    coordinate stage of the next batch starts with the step and is over before the backward begins, so only the
    forward's persistent tensor-core kernels leave it SMs) and one for loss + backward (all SMs).

        tr = PipelinedTrainer(net, example_batch, lr=1e-3)
        tr.prime(batch0)
        out0 = tr.step(batch1)        # trains on batch0; batch1 is staged and its coordinate stage started
        out1 = tr.step(batch2) ...

    Same parameters after every step as Trainer.step on the same batches in the same order (same CPU generator stream
    for the first FPS centroids).  step(None) re-uses the batch already resident in the next slot."""

    def __init__(self, net, example: Dict[str, Tensor], geometry_sms: Optional[int] = None, auto_rebuild: bool = True,
                 **kw):
        super().__init__(net, **kw)
        from . import BATCH_KEYS
        from .graph import StartRing
        dev = self.flat_param.device
        self.device = dev
        self.keys = BATCH_KEYS
        B, N, _ = example["pcs"].shape
        self.B, self.N = B, N
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.geometry_sms = min(B, sms // 4) if geometry_sms is None else int(geometry_sms)
        self.feature_sms = sms - self.geometry_sms
        self.auto_rebuild = auto_rebuild
        self.static = [{k: torch.empty_like(example[k], device=dev).copy_(example[k]) for k in BATCH_KEYS}
                       for _ in range(2)]
        self.geo = [pipeline.Geometry.empty(net, B, N, dev) for _ in range(2)]
        self.geo_stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self.geo_stream):
            self.starts = [StartRing(B, (N, net.sa1.npoint), dev) for _ in range(2)]
        self.geo_done = [torch.cuda.Event() for _ in range(2)]
        self.labels_done = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.cur = 0
        self.primed = False
        self.captures = 0
        self._capture()

    def _geometry(self, s: int):
        return pipeline.geometry_forward(self.net, self.static[s]["pcs"], self.starts[s].dev, out=self.geo[s])

    def _capture(self):
        from .graph import bn_state_key
        net, dev = self.net, self.device
        self._key = bn_state_key(net)
        buffers = {k: v.clone() for k, v in net.named_buffers()}
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():           # eager warm-up outside the capture
            for s in range(2):
                self.starts[s].draw()
                self._geometry(s)
                tape, X_raw, W_raw = Trainer._forward(self, self.static[s]["pcs"], geo=self.geo[s])
                Trainer._loss_backward(self, self.static[s], tape, X_raw, W_raw)
                del tape, X_raw, W_raw
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.g_geo, self.g_fwd, self.g_bwd, self._tapes, self.out = [], [], [], [], []
        pool = None
        for s in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                self._geometry(s)
            self.g_geo.append(g)
            # the main-stream graphs replay strictly in order on one stream: they can share one memory pool
            prev = ops.set_sm_budget(self.feature_sms)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), (torch.cuda.graph(g) if pool is None else torch.cuda.graph(g, pool=pool)):
                    tape, X_raw, W_raw = Trainer._forward(self, self.static[s]["pcs"], geo=self.geo[s])
            finally:
                ops.set_sm_budget(prev)
            pool = g.pool() if pool is None else pool
            self.g_fwd.append(g)
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g, pool=pool):
                out = Trainer._loss_backward(self, self.static[s], tape, X_raw, W_raw)
            self.g_bwd.append(g)
            self._tapes.append(tape)                       # keeps the forward's activations alive for the replays
            self.out.append(out)
        with torch.no_grad():
            for k, v in net.named_buffers():
                v.copy_(buffers[k])
        self.flat_grad.zero_()
        torch.cuda.synchronize(dev)
        self.captures += 1

    def stale(self) -> bool:
        from .graph import bn_state_key
        return self._key != bn_state_key(self.net)

    def rebuild_if_stale(self) -> bool:
        """Re-capture when mode / BatchNorm momentum changed since the last capture (update_momentum in the training
        scripts, train_Point2Cyl_without_sketch.py:357-366).  Returns True if it did."""
        if not self.stale():
            return False
        torch.cuda.synchronize(self.device)
        self._capture()
        return True

    def _stage(self, s: int, batch: Optional[Dict[str, Tensor]]):
        """Slot s <- batch (H2D when it lives in host memory), then its coordinate stage, all off the main stream."""
        cur = torch.cuda.current_stream(self.device)
        if batch is not None:
            for k in self.keys:
                if tuple(batch[k].shape) != tuple(self.static[s][k].shape):
                    raise _lib.P2CError(f"PipelinedTrainer was captured for {k} of shape "
                                        f"{tuple(self.static[s][k].shape)}, got {tuple(batch[k].shape)}: build a new one")
        self.geo_stream.wait_event(self.consumed[s])           # the slot's previous batch has been trained on
        self.geo_stream.wait_stream(cur)
        with torch.cuda.stream(self.geo_stream):
            if batch is not None:
                self.static[s]["pcs"].copy_(batch["pcs"], non_blocking=True)
            self.starts[s].draw()
            self.g_geo[s].replay()
            self.geo_done[s].record(self.geo_stream)
        self.copy_stream.wait_event(self.consumed[s])
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            if batch is not None:
                for k in self.keys:
                    if k != "pcs":
                        self.static[s][k].copy_(batch[k], non_blocking=True)
            self.labels_done[s].record(self.copy_stream)

    def prime(self, batch: Optional[Dict[str, Tensor]] = None) -> None:
        """Fill the pipeline: stage `batch` (None = the batch already resident) as the CURRENT batch."""
        if self.stale() and self.auto_rebuild:
            self.rebuild_if_stale()
        self._stage(self.cur, batch)
        self.primed = True

    @torch.no_grad()
    def step(self, next_batch: Optional[Dict[str, Tensor]] = None, fps_start=None) -> Dict[str, Tensor]:
        """One training step on the CURRENT batch; `next_batch` is staged (and its coordinate stage started) to become
        the current one.  Returns the loss dict of the batch trained on."""
        if fps_start is not None:
            raise _lib.P2CError("PipelinedTrainer draws the first FPS centroids itself (one batch ahead)")
        if self.stale():
            if not self.auto_rebuild:
                raise RuntimeError("PipelinedTrainer: mode or BatchNorm momentum changed; call rebuild_if_stale()")
            self.rebuild_if_stale()
        if not self.primed:
            self.prime(None)
        c, n = self.cur, self.cur ^ 1
        main = torch.cuda.current_stream(self.device)
        self._stage(n, next_batch)
        main.wait_event(self.geo_done[c])
        self.g_fwd[c].replay()
        main.wait_event(self.labels_done[c])
        self.g_bwd[c].replay()
        self.consumed[c].record(main)
        world = self.world()
        if world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        self.step_count += 1
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.step_count,
                      self.betas, self.eps, self.weight_decay, grad_scale=1.0 / world)
        self.cur = n
        return self.out[c]

    def join(self) -> None:
        """Make the caller's stream wait for the side-stream work of the last step() (a timed region that must contain
        ALL work launched in it ends with this)."""
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.geo_done[self.cur])
        main.wait_event(self.labels_done[self.cur])
