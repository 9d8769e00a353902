"""CUDA-graph execution of the two-stream detector (inference).

At batch 1 the forward is ~110 short kernels; launching them one by one from Python is 10x slower than the
kernels themselves.  :class:`GraphedDetector` captures one ``Model`` forward (all libicaf_b200 launches on the
capture stream, intermediate buffers in the graph's private pool) and replays it per step.  Inputs are staged
through static device buffers: device tensors are copied in with one D2D memcpy, host tensors (ideally pinned)
with one async H2D memcpy each; `uint8` frames are scaled by 1/255 inside the packing kernel, like the
reference's ``img.half() / 255`` staging (detect_twostream.py:70-80).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from .yolo_test import Model


class GraphedDetector:
    def __init__(self, model: Model, batch: int, height: int, width: int, in_dtype: torch.dtype = torch.float16,
                 device: Optional[torch.device] = None, warmup: int = 2):
        if model.training:
            raise ValueError("GraphedDetector needs model.eval()")
        self.model = model
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedDetector needs the model on a CUDA device")
        self.shape = (batch, 3, height, width)
        self.rgb = torch.zeros(self.shape, dtype=in_dtype, device=self.device)
        self.ir = torch.zeros(self.shape, dtype=in_dtype, device=self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.launches_per_step = 0
        with torch.no_grad(), torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):            # packs filters, configures kernels, warms the allocator
                self.model(self.rgb, self.ir)
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.z, self.logits, self.xs = self.model(self.rgb, self.ir)
            self.launches_per_step = ops.launch_count() - n0
        self.stream.synchronize()
        self._z_host = torch.empty(self.z.shape, dtype=self.z.dtype, pin_memory=True)

    # -- device-resident path ------------------------------------------------------------------
    def replay(self):
        """Re-run the captured forward on whatever the static input buffers hold (current stream)."""
        self.graph.replay()
        return self.z, self.logits, self.xs

    def __call__(self, rgb: torch.Tensor, ir: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, list]:
        """The reference-facing call: ``pred = model(img_rgb, img_ir)``.  Device or host inputs of the captured
        shape / dtype.  Returns device tensors (z, logits, [x0,x1,x2]) valid until the next call."""
        if tuple(rgb.shape) != self.shape or tuple(ir.shape) != self.shape:
            raise ValueError(f"GraphedDetector was captured for {self.shape}, got {tuple(rgb.shape)}")
        self.rgb.copy_(rgb, non_blocking=True)
        self.ir.copy_(ir, non_blocking=True)
        self.graph.replay()
        return self.z, self.logits, self.xs

    def infer_to_host(self, rgb_host: torch.Tensor, ir_host: torch.Tensor) -> torch.Tensor:
        """End-to-end step from (pinned) host frames to the decoded predictions on the host: H2D, forward, D2H, sync."""
        self(rgb_host, ir_host)
        self._z_host.copy_(self.z, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._z_host
