"""Host <-> device streaming around `DPTDepthModel.forward` for batched inference from host memory.

`StreamingPredictor.run` overlaps the pinned-host -> device copy of batch i+1 and the device -> host
read-back of batch i-1 with the forward pass of batch i (two input slots, three CUDA streams, events
only — no host synchronisation inside the loop).  This is the public end-to-end entry point measured
by bench.py's `e2e` leg; the reference has no counterpart (demo.py moves one image at a time,
demo.py:132,147).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch


class StreamingPredictor:
    """`model`: a DPTDepthModel, or any callable mapping a device batch to one tensor or a list of tensors (e.g. the
    depth and the normal network on the same images); `host_outputs[i % len]` is then a pinned tensor or a matching
    list of pinned tensors."""

    def __init__(self, model, device: torch.device):
        self.model = model
        self.device = device
        self.copy_in = torch.cuda.Stream(device)
        self.copy_out = torch.cuda.Stream(device)
        self._slots: List[torch.Tensor] = []

    def _slot(self, k: int, like: torch.Tensor) -> torch.Tensor:
        while len(self._slots) <= k:
            self._slots.append(torch.empty(0, device=self.device))
        s = self._slots[k]
        if s.shape != like.shape or s.dtype != like.dtype:
            s = self._slots[k] = torch.empty(like.shape, dtype=like.dtype, device=self.device)
        return s

    @torch.no_grad()
    def run(self, host_inputs: Iterable[torch.Tensor], host_outputs: Sequence[torch.Tensor]) -> int:
        """host_inputs: pinned [B,3,H,W] float tensors; host_outputs[i % len] (pinned) receives result i.
        Returns the number of batches processed.  Work is enqueued on the current stream + two copy
        streams; on return the CURRENT stream waits for the last read-back."""
        main = torch.cuda.current_stream(self.device)
        consumed = [None, None]
        n = 0
        last_out = None
        for i, h in enumerate(host_inputs):
            k = i & 1
            slot = self._slot(k, h)
            with torch.cuda.stream(self.copy_in):
                if consumed[k] is not None:
                    self.copy_in.wait_event(consumed[k])       # forward i-2 has read this slot
                slot.copy_(h, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_in)
            main.wait_event(ready)
            y = self.model(slot)
            done = torch.cuda.Event()
            done.record(main)
            consumed[k] = done
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(done)
                dst = host_outputs[i % len(host_outputs)]
                if torch.is_tensor(y):
                    ys, dsts = [y], [dst]
                else:
                    ys, dsts = list(y), list(dst)
                for yy, dd in zip(ys, dsts):
                    dd.copy_(yy, non_blocking=True)
                    yy.record_stream(self.copy_out)
                last_out = torch.cuda.Event()
                last_out.record(self.copy_out)
            n += 1
        if last_out is not None:
            main.wait_event(last_out)
        return n
