"""Double-buffered host <-> device pipeline around HiT_SIR.forward (the caller side of experiments/experiment.py:736-743,
599-600: `.to(device)` -> model -> `.cpu()`), so that the PCIe copies of batch i overlap the kernels of batch i-1 / i+1.

    pipe = HostPipeline(model, "cuda:0")
    for x_host, y_host in batches:          # pinned CPU tensors
        pipe.submit(x_host, y_host)
    pipe.wait()                              # every y_host is complete

Three CUDA streams: H2D copies, the model's compute stream (the current stream), D2H copies; events order them per slot.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class HostPipeline:
    def __init__(self, model, device, depth: int = 2):
        self.model = model
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device: hitsir_b200 has no CPU path.")
        self.depth = depth
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._slots: List[dict] = [dict(dx=None, done=None) for _ in range(depth)]
        self._i = 0

    def submit(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        slot = self._slots[self._i % self.depth]
        self._i += 1
        if slot["done"] is not None:
            slot["done"].synchronize()                     # the slot's previous result has reached the host
        compute = torch.cuda.current_stream(self.device)
        if slot["dx"] is None or slot["dx"].shape != x_host.shape:
            # the block may come back from the caching allocator with work of its previous user still queued on the compute stream
            # (the allocator only orders reuse on the allocating stream): allocate on the copy stream and make it wait for compute
            self.h2d.wait_stream(compute)
            with torch.cuda.stream(self.h2d):
                slot["dx"] = torch.empty(x_host.shape, dtype=torch.float32, device=self.device)
            slot["dx"].record_stream(compute)
        ev_in, ev_out, ev_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(self.h2d):
            slot["dx"].copy_(x_host, non_blocking=True)
            ev_in.record(self.h2d)
        compute.wait_event(ev_in)
        with torch.no_grad():
            y = self.model(slot["dx"])
        ev_out.record(compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(ev_out)
            y_host.copy_(y, non_blocking=True)
            ev_done.record(self.d2h)
        y.record_stream(self.d2h)
        slot["done"] = ev_done

    def wait(self) -> None:
        for slot in self._slots:
            if slot["done"] is not None:
                slot["done"].synchronize()
