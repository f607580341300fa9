"""A HARD partition of the GPU's SMs between concurrent stages: CUDA green contexts (driver API, CUDA >= 12.4).

The pipelined forward+loss step runs three stages at once (graph.DeepPipelinedForwardLoss).  With plain streams the
split is soft: the persistent layer kernels are launched with 116 CTAs and the coordinate / loss kernels take whatever
SMs are free - including, at every layer-kernel boundary, SMs the next layer kernel is about to need (hundreds of small
CTAs flood the freed SMs; CTAs of a statically strided persistent grid that start late set the kernel's time).  A green
context owns a fixed set of SMs: kernels launched into its streams - also as nodes of a CUDA graph captured on such a
stream, whichever stream replays it (measured, tools/green_probe.py) - run on those SMs only.

`SmPartition(device, small)` splits the device into a group of `small` SMs (a multiple of 8 on sm_90+) and the rest and
hands out torch streams of either side.  Needs the `cuda-python` driver bindings; `SmPartition.create` returns None when
they (or the API) are missing, and the caller falls back to the soft split.
"""
from __future__ import annotations

from typing import Optional

import torch


class SmPartition:
    def __init__(self, cu, dev, ctx_small, ctx_big, n_small: int, n_big: int):
        self._cu, self._dev = cu, dev
        self._ctx = (ctx_small, ctx_big)           # keep the green contexts alive as long as their streams are used
        self.small_sms, self.big_sms = n_small, n_big
        self._streams = []

    @staticmethod
    def create(device: torch.device, small_sms: int) -> Optional["SmPartition"]:
        try:
            from cuda.bindings import driver as cu
        except Exception:
            try:
                from cuda import cuda as cu          # older cuda-python layout
            except Exception:
                return None
        ok = cu.CUresult.CUDA_SUCCESS
        try:
            torch.zeros(1, device=device)            # the primary context exists and is current
            err, dev = cu.cuDeviceGet(device.index or 0)
            if err != ok:
                return None
            err, res = cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM)
            if err != ok or small_sms <= 0 or small_sms >= res.sm.smCount:
                return None
            err, groups, nb, remaining = cu.cuDevSmResourceSplitByCount(1, res, 0, int(small_sms))
            if err != ok or nb < 1 or remaining.sm.smCount == 0:
                return None
            ctxs = []
            for r in (groups[0], remaining):
                err, desc = cu.cuDevResourceGenerateDesc([r], 1)
                if err != ok:
                    return None
                err, g = cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM)
                if err != ok:
                    return None
                ctxs.append(g)
            return SmPartition(cu, dev, ctxs[0], ctxs[1], int(groups[0].sm.smCount), int(remaining.sm.smCount))
        except Exception:
            return None

    def _stream(self, which: int) -> torch.cuda.Stream:
        cu = self._cu
        err, s = cu.cuGreenCtxStreamCreate(self._ctx[which], cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0)
        if err != cu.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"cuGreenCtxStreamCreate failed: {err}")
        self._streams.append(s)
        return torch.cuda.ExternalStream(int(s))

    def small_stream(self) -> torch.cuda.Stream:
        """A new stream whose kernels run on the small group of SMs."""
        return self._stream(0)

    def big_stream(self) -> torch.cuda.Stream:
        """A new stream whose kernels run on the remaining SMs."""
        return self._stream(1)
