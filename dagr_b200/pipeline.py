"""Host -> device input pipeline: the next batch's H2D copy runs on a side stream while the current batch is
processed (the reference copies synchronously inside the loop: utils/testing.py:29)."""
from __future__ import annotations

import torch


class Prefetcher:
    """iterates over pinned host batches; `next()` returns a device batch whose copy was enqueued one step earlier."""

    def __init__(self, batches, device, transform=None):
        self.batches, self.device, self.transform = list(batches), device, transform
        self.stream = torch.cuda.Stream(device=device)
        self._next = None
        self._i = 0

    def _issue(self, i):
        with torch.cuda.stream(self.stream):
            d = self.batches[i % len(self.batches)].to(self.device, non_blocking=True)
            if self.transform is not None:
                d = self.transform(d)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return d, ev

    def next(self):
        if self._next is None:
            self._next = self._issue(self._i)
        d, ev = self._next
        self._i += 1
        self._next = self._issue(self._i)                   # overlaps with the caller's compute on `d`
        torch.cuda.current_stream().wait_event(ev)
        for v in d.__dict__.values():                       # tensors were allocated on the side stream
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(torch.cuda.current_stream())
        return d
