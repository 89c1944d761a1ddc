"""CUDA-graph replay of the forward pass for one fixed input shape.

A HiT-SIR-pro forward is 263 kernel launches.  At batch sizes where each kernel runs for a few microseconds (the reference's own
1x3x64x64 case: 6.9 ms eager) the step is bound by launch latency and the gaps between dependent kernels, not by any roofline, so the
whole launch sequence of `HiT_SIR.forward` -- tensor maps are built on the host and passed by value, the library never synchronises or
allocates -- is captured once into a `torch.cuda.CUDAGraph` and replayed.  Weights are baked into the captured launches as device
pointers into the handle's packed buffers: re-capture after changing parameters (`GraphedForward.recapture()`).
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
        self.recapture()

    def recapture(self):
        dev = self.x.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):            # packs weights, sizes the workspace, opts kernels in to their shared memory
                self.model(self.x)
        side.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side), torch.no_grad():
            self.y = self.model(self.x)
        torch.cuda.current_stream(dev).wait_stream(side)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Copies `x` into the captured input buffer, replays, returns the captured output buffer (overwritten by the next call)."""
        if x.shape != self.x.shape:
            raise RuntimeError(f"GraphedForward was captured for {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y
