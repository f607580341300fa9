"""CUDA-graph replay of one forward+loss step.

The step is ~80 short kernels; launched one by one from Python their launch latency is a visible part of
the step.  With the on-device assignment (p2c_hungarian) nothing in the step needs the host, so the whole
sequence - every libp2c kernel, the dropout-mask op and the small torch glue - is captured once into a CUDA
graph and replayed.  The only per-step host work mirrors the reference: the first FPS centroid of each level
is drawn from the CPU generator (models/pointnet_util.py:75) and copied to the device.

Re-capture is needed when anything baked into the launch arguments changes: batch shape, loss weights,
train/eval mode, or the BatchNorm momentum (train scripts change it via update_momentum).

The step is captured as TWO graphs - backbone, then loss - so that a batch arriving from host memory can overlap
its copies with the compute: only the point coordinates (3 of the 10.5 MB per step) have to be on the device before the
backbone starts; normals, labels and ground-truth axes / centres are copied on a side stream while the backbone
runs and the loss graph waits for that stream's event.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import BATCH_KEYS, _lib, pipeline


class StartRing:
    """First-centroid draws of the two FPS levels (models/pointnet_util.py:75: CPU generator, then moved) for graph
    replay.  The host may run several replays ahead of the GPU, so the pinned staging buffers form a ring and each slot
    is rewritten only after the async copy that last read it has completed (an event per slot)."""

    def __init__(self, B: int, ranges, device, depth: int = 4):
        self.B, self.ranges = B, tuple(ranges)
        self.host = [[torch.zeros(B, dtype=torch.long).pin_memory() for _ in self.ranges] for _ in range(depth)]
        self.done = [None] * depth
        self.dev = [torch.zeros(B, dtype=torch.long, device=device) for _ in self.ranges]
        self.i = 0

    def draw(self):
        """same calls, same order as the reference: sa1 draws from [0,N), then sa2 from [0,npoint1)"""
        slot = self.i % len(self.host)
        self.i += 1
        if self.done[slot] is not None:
            self.done[slot].synchronize()
        for h, d, hi in zip(self.host[slot], self.dev, self.ranges):
            h.copy_(torch.randint(0, hi, (self.B,), dtype=torch.long))
            d.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.done[slot] = ev


def bn_state_key(net):
    return (net.training,) + tuple(m.momentum for m in net.modules()
                                   if isinstance(m, torch.nn.modules.batchnorm._BatchNorm))


class GraphedForwardLoss:
    """auto_rebuild: when the train/eval mode or a BatchNorm momentum changed since capture (the training scripts call
    update_momentum, train_Point2Cyl_without_sketch.py:357-366) the graphs are re-captured on the next call instead
    of raising; `rebuild_if_stale()` does the same explicitly."""

    def __init__(self, net, example: Dict[str, torch.Tensor], weights=(1.0,) * 5, norm_eig: bool = False,
                 precision: Optional[str] = None, warmup: int = 2, auto_rebuild: bool = True):
        self.net = net
        dev = next(net.parameters()).device
        self.device = dev
        self.static = {k: torch.empty_like(example[k], device=dev).copy_(example[k]) for k in BATCH_KEYS}
        B, N, _ = self.static["pcs"].shape
        self.B, self.N, self.S1 = B, N, net.sa1.npoint
        self.starts = StartRing(B, (N, self.S1), dev)
        self.start_dev = self.starts.dev
        self.weights, self.norm_eig, self.precision = weights, norm_eig, precision
        self.auto_rebuild = auto_rebuild
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = torch.cuda.Event()
        self.captures = 0
        self._capture(warmup)

    def _capture(self, warmup: int):
        dev = self.device
        self._key = self._state_key()
        # eager warm-up outside the capture (one-time lazy init, allocator growth).  Train-mode passes update the
        # BatchNorm running statistics: they are restored afterwards, so building the graph is not a training step.
        buffers = {k: v.clone() for k, v in self.net.named_buffers()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._draw_starts()
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()            # backbone
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.X_raw, self.W_raw = pipeline.backbone_forward(self.net, self.static["pcs"], self.start_dev,
                                                               precision=self.precision)
        self.loss_graph = torch.cuda.CUDAGraph()       # loss block on the backbone's static outputs
        with torch.no_grad(), torch.cuda.graph(self.loss_graph, pool=self.graph.pool()):
            self.out = pipeline.loss_forward(self.static["pcs"], self.X_raw, self.W_raw, self.static["normals"],
                                             self.static["inst"], self.static["bb"], self.static["axes"],
                                             self.static["centers"], self.weights, self.norm_eig)
            self.out.update(X_raw=self.X_raw, W_raw=self.W_raw)
        with torch.no_grad():
            for k, v in self.net.named_buffers():
                v.copy_(buffers[k])
        self.captures += 1

    def rebuild_if_stale(self) -> bool:
        """Re-capture when mode / BatchNorm momentum changed since the last capture.  Returns True if it did."""
        if not self.stale():
            return False
        torch.cuda.synchronize(self.device)
        self._capture(warmup=1)
        return True

    def _state_key(self):
        return bn_state_key(self.net)

    def _run(self):
        return pipeline.forward_loss(self.net, self.static, fps_start=self.start_dev, weights=self.weights,
                                     norm_eig=self.norm_eig, precision=self.precision)

    def _draw_starts(self):
        self.starts.draw()

    def stale(self) -> bool:
        return self._key != self._state_key()

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """batch: host (pinned) or device tensors copied into the graph's static inputs; None = reuse them."""
        if self.stale():
            if not self.auto_rebuild:
                raise RuntimeError("GraphedForwardLoss: mode or BatchNorm momentum changed since capture; call "
                                   "rebuild_if_stale()")
            self.rebuild_if_stale()
        cur = torch.cuda.current_stream(self.device)
        if batch is not None:
            for k in BATCH_KEYS:                                           # copy_ would silently broadcast a short batch
                if tuple(batch[k].shape) != tuple(self.static[k].shape):
                    raise _lib.P2CError(f"GraphedForwardLoss was captured for {k} of shape "
                                        f"{tuple(self.static[k].shape)}, got {tuple(batch[k].shape)}: build a new one")
            self.static["pcs"].copy_(batch["pcs"], non_blocking=True)      # the backbone needs only the coordinates
            self.copy_stream.wait_stream(cur)                              # (the previous step's loss has read the labels)
            with torch.cuda.stream(self.copy_stream):
                for k in BATCH_KEYS:
                    if k != "pcs":
                        self.static[k].copy_(batch[k], non_blocking=True)
                self.copied.record(self.copy_stream)
        self._draw_starts()
        self.graph.replay()
        if batch is not None:
            cur.wait_event(self.copied)
        self.loss_graph.replay()
        return self.out


class PipelinedForwardLoss:
    """Forward+loss over a STREAM of batches as a two-stage software pipeline (one batch deep).

    The coordinate-only stage of a batch (pipeline.geometry_forward: both FPS levels, both ball queries, the 3-NN
    searches - ~0.6 ms of short, mostly serial kernels on a few SMs) needs nothing but the coordinates, so it runs for
    batch i+1 on a second stream while batch i goes through the per-point MLP layers and the loss on the main stream.
    Each stage is a CUDA graph per buffer slot (two slots, ping-pong); events order slot reuse.  The persistent
    tensor-core kernels of the feature stage are captured with `p2c_set_sm_budget(SMs - geometry_sms)`, which leaves
    the geometry kernels (one FPS CTA per cloud) SMs of their own.

        pipe = PipelinedForwardLoss(net, example_batch)
        pipe.prime(batch0)                       # fills the pipeline: copies batch0 in, starts its geometry
        out0 = pipe.step(batch1)                 # loss of batch0; batch1 is staged and its geometry started
        out1 = pipe.step(batch2) ...

    Every step() does one batch's worth of every kind of work (H2D of one batch, one geometry stage, one feature
    stage + loss); results are identical to pipeline.forward_loss on the same batches in the same order (same CPU
    generator stream for the first FPS centroids).  step(None) re-uses the batch already resident in the next slot.
    """

    def __init__(self, net, example: Dict[str, torch.Tensor], weights=(1.0,) * 5, norm_eig: bool = False,
                 precision: Optional[str] = None, geometry_sms: Optional[int] = None, auto_rebuild: bool = True):
        self.net = net
        dev = next(net.parameters()).device
        self.device = dev
        self.weights, self.norm_eig, self.precision, self.auto_rebuild = weights, norm_eig, precision, auto_rebuild
        B, N, _ = example["pcs"].shape
        self.B, self.N = B, N
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.geometry_sms = min(B, sms // 4) if geometry_sms is None else int(geometry_sms)
        self.feature_sms = sms - self.geometry_sms
        self.static = [{k: torch.empty_like(example[k], device=dev).copy_(example[k]) for k in BATCH_KEYS}
                       for _ in range(2)]
        self.geo = [pipeline.Geometry.empty(net, B, N, dev) for _ in range(2)]
        self.geo_stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self.geo_stream):
            self.starts = [StartRing(B, (N, net.sa1.npoint), dev) for _ in range(2)]
        self.geo_done = [torch.cuda.Event() for _ in range(2)]
        self.labels_done = [torch.cuda.Event() for _ in range(2)]
        self.feat_done = [torch.cuda.Event() for _ in range(2)]
        self.cur = 0
        self.primed = False
        self.captures = 0
        self._capture()

    # ---- capture -------------------------------------------------------------------------------------------------
    def _geometry(self, s: int):
        return pipeline.geometry_forward(self.net, self.static[s]["pcs"], self.starts[s].dev, out=self.geo[s])

    def _features(self, s: int):
        b = self.static[s]
        out = pipeline.forward_loss(self.net, b, weights=self.weights, norm_eig=self.norm_eig,
                                    precision=self.precision, geo=self.geo[s])
        return out

    def _capture(self):
        from . import ops
        dev = self.device
        self._key = bn_state_key(self.net)
        buffers = {k: v.clone() for k, v in self.net.named_buffers()}
        torch.cuda.synchronize(dev)
        prev = ops.set_sm_budget(self.feature_sms)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():       # eager warm-up outside the capture
                for s in range(2):
                    self.starts[s].draw()
                    self._geometry(s)
                    self._features(s)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.g_geo, self.g_feat, self.out = [], [], []
            for s in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g):
                    self._geometry(s)
                self.g_geo.append(g)
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g):
                    out = self._features(s)
                self.g_feat.append(g)
                self.out.append(out)
        finally:
            ops.set_sm_budget(prev)
        with torch.no_grad():
            for k, v in self.net.named_buffers():
                v.copy_(buffers[k])
        torch.cuda.synchronize(dev)
        self.captures += 1

    def stale(self) -> bool:
        return self._key != bn_state_key(self.net)

    def rebuild_if_stale(self) -> bool:
        if not self.stale():
            return False
        torch.cuda.synchronize(self.device)
        self._capture()
        return True

    # ---- running -------------------------------------------------------------------------------------------------
    def _check(self, batch):
        for k in BATCH_KEYS:
            if tuple(batch[k].shape) != tuple(self.static[0][k].shape):
                raise _lib.P2CError(f"PipelinedForwardLoss was captured for {k} of shape "
                                    f"{tuple(self.static[0][k].shape)}, got {tuple(batch[k].shape)}: build a new one")

    def _stage(self, s: int, batch: Optional[Dict[str, torch.Tensor]]):
        """Slot s <- batch (H2D when it lives in host memory), then its geometry stage, all off the main stream."""
        cur = torch.cuda.current_stream(self.device)
        if batch is not None:
            self._check(batch)
        self.geo_stream.wait_event(self.feat_done[s])              # the slot's previous batch has been consumed
        self.geo_stream.wait_stream(cur)                           # (and whatever produced `batch` on the caller's stream)
        with torch.cuda.stream(self.geo_stream):
            if batch is not None:
                self.static[s]["pcs"].copy_(batch["pcs"], non_blocking=True)
            self.starts[s].draw()
            self.g_geo[s].replay()
            self.geo_done[s].record(self.geo_stream)
        self.copy_stream.wait_event(self.feat_done[s])
        self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            if batch is not None:
                for k in BATCH_KEYS:
                    if k != "pcs":
                        self.static[s][k].copy_(batch[k], non_blocking=True)
            self.labels_done[s].record(self.copy_stream)

    def prime(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """Fill the pipeline: stage `batch` (None = the batch already resident) as the CURRENT batch."""
        if self.stale() and self.auto_rebuild:
            self.rebuild_if_stale()
        self._stage(self.cur, batch)
        self.primed = True

    def step(self, next_batch: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Loss of the current batch; `next_batch` is staged (and its geometry started) to become the current one."""
        if self.stale():
            if not self.auto_rebuild:
                raise RuntimeError("PipelinedForwardLoss: mode or BatchNorm momentum changed; call rebuild_if_stale()")
            self.rebuild_if_stale()
        if not self.primed:
            self.prime(None)
        c, n = self.cur, self.cur ^ 1
        main = torch.cuda.current_stream(self.device)
        self._stage(n, next_batch)
        main.wait_event(self.geo_done[c])
        main.wait_event(self.labels_done[c])
        self.g_feat[c].replay()
        self.feat_done[c].record(main)
        self.cur = n
        return self.out[c]

    def join(self) -> None:
        """Make the caller's stream wait for the side-stream work of the last step() (the staging and geometry of the
        batch that is now current).  Not needed for correctness - step() orders everything it uses - but a timed
        region that must contain ALL work launched in it ends with this."""
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.geo_done[self.cur])
        main.wait_event(self.labels_done[self.cur])


class DeepPipelinedForwardLoss:
    """PipelinedForwardLoss with the LOSS BLOCK as a third stage: per step

        geometry of batch i+1 (second stream)  |  backbone layers of batch i (main stream)  |  loss of batch i-1 (third stream)

    over three buffer slots.  The loss block is ~0.14 ms of short, dependent kernels on a handful of SMs (statistics pass,
    cost, on-device assignment, base/barrel pass, 3x3 eigen-solves); at the end of the main stream it keeps the feature
    stage's SMs idle - beside the next batch's layers it is free.  Every step() still does one batch's worth of every
    kind of work, launched and (after join()) completed inside the step; what it returns is the loss of the batch
    staged TWO calls earlier (None while the pipeline fills; `flush()` drains the last one).  Same results as
    pipeline.forward_loss on the same batches in the same order.

    partition="soft" (default): plain streams and a capped grid for the persistent layer kernels.  partition="green": the
    SMs are split HARD with CUDA green contexts (point2cyl_b200.partition; needs the cuda-python driver bindings, falls back
    to "soft" without them) - the layers' graph is captured on a stream that owns all but `geometry_sms` SMs (a multiple of
    8, default 48), the coordinate and loss graphs on streams that own the rest: neither side's CTAs can land on the other's
    SMs.  Measured at config 2: device-resident step 1.132 -> 1.096 ms at 48 SMs (32: 1.39, 40: 1.20, 56: 1.17), but the
    end-to-end step (host copies in front of the coordinate stage, which can no longer borrow SMs) 1.135 -> 1.176 ms - hence
    not the default."""

    SLOTS = 3

    def __init__(self, net, example: Dict[str, torch.Tensor], weights=(1.0,) * 5, norm_eig: bool = False,
                 precision: Optional[str] = None, geometry_sms: Optional[int] = None, auto_rebuild: bool = True,
                 partition: str = "soft"):
        self.net = net
        dev = next(net.parameters()).device
        self.device = dev
        self.weights, self.norm_eig, self.precision, self.auto_rebuild = weights, norm_eig, precision, auto_rebuild
        B, N, _ = example["pcs"].shape
        self.B, self.N = B, N
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.part = None
        if partition == "green":
            from .partition import SmPartition
            want = 48 if geometry_sms is None else int(geometry_sms)
            self.part = SmPartition.create(dev, (want + 7) // 8 * 8)
        if self.part is not None:
            self.geometry_sms, self.feature_sms = self.part.small_sms, self.part.big_sms
        else:
            self.geometry_sms = min(B, sms // 4) if geometry_sms is None else int(geometry_sms)
            self.feature_sms = sms - self.geometry_sms
        self.partition = "green" if self.part is not None else "soft"
        S = self.SLOTS
        self.static = [{k: torch.empty_like(example[k], device=dev).copy_(example[k]) for k in BATCH_KEYS} for _ in range(S)]
        self.geo = [pipeline.Geometry.empty(net, B, N, dev) for _ in range(S)]
        self.geo_stream = self.part.small_stream() if self.part is not None else torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.loss_stream = self.part.small_stream() if self.part is not None else torch.cuda.Stream(device=dev)
        self.feat_stream = self.part.big_stream() if self.part is not None else None
        with torch.cuda.stream(self.geo_stream):
            self.starts = [StartRing(B, (N, net.sa1.npoint), dev) for _ in range(S)]
        ev = lambda: [torch.cuda.Event() for _ in range(S)]
        self.geo_done, self.labels_done, self.back_done, self.loss_done = ev(), ev(), ev(), ev()
        self.i = 0                    # batch index of the CURRENT batch (slot i % 3)
        self.staged = -1              # highest batch index staged
        self.backboned = -1           # highest batch index whose backbone has been launched
        self.lossed = -1
        self.captures = 0
        self._capture()

    def _geometry(self, s: int):
        return pipeline.geometry_forward(self.net, self.static[s]["pcs"], self.starts[s].dev, out=self.geo[s])

    def _capture(self):
        from . import ops
        dev = self.device
        self._key = bn_state_key(self.net)
        buffers = {k: v.clone() for k, v in self.net.named_buffers()}
        torch.cuda.synchronize(dev)
        prev = ops.set_sm_budget(self.feature_sms)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():       # eager warm-up outside the capture
                for s in range(self.SLOTS):
                    self.starts[s].draw()
                    self._geometry(s)
                    pipeline.forward_loss(self.net, self.static[s], weights=self.weights, norm_eig=self.norm_eig,
                                          precision=self.precision, geo=self.geo[s])
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.g_geo, self.g_back, self.g_loss, self.out = [], [], [], []
            for s in range(self.SLOTS):
                b = self.static[s]
                # a kernel node keeps the green context of the stream it was captured on: the capture streams ARE the partition
                cap_small = self.geo_stream if self.part is not None else None
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, stream=cap_small):
                    self._geometry(s)
                self.g_geo.append(g)
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, stream=self.feat_stream):
                    X_raw, W_raw = pipeline.backbone_forward(self.net, b["pcs"], None, precision=self.precision,
                                                             geo=self.geo[s])
                self.g_back.append(g)
                gl = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(gl, pool=g.pool(), stream=self.loss_stream if self.part is not None else None):
                    out = pipeline.loss_forward(b["pcs"], X_raw, W_raw, b["normals"], b["inst"], b["bb"], b["axes"],
                                                b["centers"], self.weights, self.norm_eig)
                    out.update(X_raw=X_raw, W_raw=W_raw)
                self.g_loss.append(gl)
                self.out.append(out)
        finally:
            ops.set_sm_budget(prev)
        with torch.no_grad():
            for k, v in self.net.named_buffers():
                v.copy_(buffers[k])
        torch.cuda.synchronize(dev)
        self.captures += 1

    def stale(self) -> bool:
        return self._key != bn_state_key(self.net)

    def rebuild_if_stale(self) -> bool:
        if not self.stale():
            return False
        torch.cuda.synchronize(self.device)
        self._capture()
        return True

    def _stage(self, j: int, batch: Optional[Dict[str, torch.Tensor]], after: Optional[torch.cuda.Event] = None):
        """Batch j -> slot j % 3 (H2D when it lives in host memory), then its geometry stage, all off the main stream.
        after: the point of the caller's stream the staging is ordered behind (default: everything enqueued so far)."""
        s = j % self.SLOTS
        cur = torch.cuda.current_stream(self.device)
        if batch is not None:
            for k in BATCH_KEYS:
                if tuple(batch[k].shape) != tuple(self.static[0][k].shape):
                    raise _lib.P2CError(f"DeepPipelinedForwardLoss was captured for {k} of shape "
                                        f"{tuple(self.static[0][k].shape)}, got {tuple(batch[k].shape)}: build a new one")
        for st in (self.geo_stream, self.copy_stream):
            st.wait_event(self.back_done[s])                       # the slot's previous batch (j - 3) has been consumed:
            st.wait_event(self.loss_done[s])                       # its backbone read the geometry, its loss the labels / pcs
            if after is not None:
                st.wait_event(after)
            else:
                st.wait_stream(cur)
        with torch.cuda.stream(self.geo_stream):
            if batch is not None:
                self.static[s]["pcs"].copy_(batch["pcs"], non_blocking=True)
            self.starts[s].draw()
            self.g_geo[s].replay()
            self.geo_done[s].record(self.geo_stream)
        with torch.cuda.stream(self.copy_stream):
            if batch is not None:
                for k in BATCH_KEYS:
                    if k != "pcs":
                        self.static[s][k].copy_(batch[k], non_blocking=True)
            self.labels_done[s].record(self.copy_stream)
        self.staged = j

    def prime(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> None:
        if self.stale() and self.auto_rebuild:
            self.rebuild_if_stale()
        self._stage(self.i, batch)

    def _launch_loss(self, j: int):
        s = j % self.SLOTS
        self.loss_stream.wait_event(self.back_done[s])
        self.loss_stream.wait_event(self.labels_done[s])
        with torch.cuda.stream(self.loss_stream):
            self.g_loss[s].replay()
            self.loss_done[s].record(self.loss_stream)
        self.lossed = j

    def step(self, next_batch: Optional[Dict[str, torch.Tensor]] = None) -> Optional[Dict[str, torch.Tensor]]:
        """Backbone of the current batch i, loss of batch i-1 beside it, `next_batch` staged as batch i+1.  Returns the
        outputs of batch i-1 (None on the first call after prime())."""
        if self.stale():
            if not self.auto_rebuild:
                raise RuntimeError("DeepPipelinedForwardLoss: mode or BatchNorm momentum changed; call rebuild_if_stale()")
            self.rebuild_if_stale()
        if self.staged < self.i:
            self.prime(None)
        i = self.i
        c = i % self.SLOTS
        main = torch.cuda.current_stream(self.device)
        # enqueue order = placement order on an idle GPU: the layers first (their persistent CTAs take their SMs before the
        # small CTAs of the other two stages can), then the loss block, then the staging + coordinate stage of the next batch
        entry = torch.cuda.Event()
        entry.record(main)                               # whatever the caller enqueued before this step
        fs = self.feat_stream if self.feat_stream is not None else main
        if fs is not main:
            fs.wait_event(entry)
        fs.wait_event(self.geo_done[c])
        fs.wait_event(self.loss_done[c])                 # the loss of batch i - 3 has read this slot's network outputs
        with torch.cuda.stream(fs):
            self.g_back[c].replay()
            self.back_done[c].record(fs)
        prev = None
        if self.backboned == i - 1 and i - 1 >= 0 and self.lossed < i - 1:
            self._launch_loss(i - 1)                     # beside the layers of batch i
            prev = (i - 1) % self.SLOTS
        self._stage(i + 1, next_batch, after=entry)      # (ordered behind the step's entry, not behind its layers)
        if fs is not main:
            main.wait_event(self.back_done[c])           # the caller's stream stays ordered behind the layers
        self.backboned = i
        self.i = i + 1
        if prev is None:
            return None
        main.wait_event(self.loss_done[prev])            # the caller's stream may read the returned tensors
        return self.out[prev]

    def flush(self) -> Optional[Dict[str, torch.Tensor]]:
        """Loss of the last batch whose backbone has run (drains the pipeline's third stage)."""
        j = self.backboned
        if j < 0:
            return None
        if self.lossed < j:
            self._launch_loss(j)
        torch.cuda.current_stream(self.device).wait_event(self.loss_done[j % self.SLOTS])
        return self.out[j % self.SLOTS]

    def join(self) -> None:
        """Make the caller's stream wait for everything launched by the last step() on the side streams."""
        main = torch.cuda.current_stream(self.device)
        s = self.i % self.SLOTS
        main.wait_event(self.geo_done[s])
        main.wait_event(self.labels_done[s])
        if self.lossed >= 0:
            main.wait_event(self.loss_done[self.lossed % self.SLOTS])
