"""CUDA-graph replay of the forward pass for one fixed input shape.

A HiT-SIR-pro forward is ~260 kernel launches.  At batch sizes where each kernel runs for a few microseconds (the reference's own
1x3x64x64 case) the step is bound by launch latency and the gaps between dependent kernels, not by any roofline, so the whole launch
sequence of `HiT_SIR.forward` -- tensor maps are built on the host and passed by value, the library never synchronises or allocates --
is captured once into a `torch.cuda.CUDAGraph` and replayed.

The captured launches hold raw device pointers into (a) the scratch workspace and (b) the handle's packed-weight buffers.  The graph
therefore OWNS its workspace (allocated here, never the module's cached one, which an eager forward with another shape would free),
and it records the module's weights key and native handle at capture time: `__call__` re-captures by itself when either has changed
(`load_state_dict`, an in-place weight update, `.to()`), because re-packing frees the buffers the old graph points to.
"""
from __future__ import annotations

import torch


class GraphedForward:
    def __init__(self, model, example: torch.Tensor, warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA example input: there is no CPU path.")
        self.model = model
        self.x = example.detach().to(torch.float32).contiguous().clone()
        self.warmup = warmup
        self.graph = None
        self.y = None
        self.workspace = None
        self.captures = 0
        self._key = None
        self.recapture()

    def _state_key(self):
        return (self.model._weights_key(), self.model.native_handle(self.x.device))

    def recapture(self):
        dev = self.x.device
        B, _, H, W = self.x.shape
        self.graph = None                               # drop the old graph before its workspace
        model = self.model
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            model(self.x)                               # creates the handle and packs the weights (allocates: must precede the capture)
            self.workspace = model.new_workspace(dev, B, H, W)
            model._forced_workspace = self.workspace
            try:
                for _ in range(self.warmup):            # opts kernels in to their shared memory on this device
                    model(self.x)
                side.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    self.y = model(self.x)
            finally:
                model._forced_workspace = None
        torch.cuda.current_stream(dev).wait_stream(side)
        self._key = self._state_key()
        self.captures += 1

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Copies `x` into the captured input buffer, replays, returns the captured output buffer (overwritten by the next call)."""
        if x.shape != self.x.shape:
            raise RuntimeError(f"GraphedForward was captured for {tuple(self.x.shape)}, got {tuple(x.shape)}")
        if self._state_key() != self._key:              # weights re-packed or handle replaced: the old launches point at freed buffers
            self.recapture()
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y
