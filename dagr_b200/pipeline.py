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


class PipelinedDetector:
    """Throughput-oriented front end of `DAGR.forward` for independent batches (reset=True forwards).

        det = PipelinedDetector(model)
        h = det.submit(batch0)
        for batch in batches[1:]:
            h_next = det.submit(batch)        # enqueue step i+1 ...
            detections = h.result()           # ... before blocking on step i
            h = h_next

    `submit` enqueues the event-level kernels on the caller's stream and the coarse stack + NMS + the device->host
    copy of the detections on the engine's side stream (`Engine.overlap`), so consecutive steps overlap on the GPU and the
    host never waits inside `submit`.  `result()` blocks on that step's event and returns the same list of dicts
    (`boxes`, `scores`, `labels`) as `DAGR.forward(...)[0]` (src/dagr/model/utils.py:104-108), as host tensors.
    At most two steps may be in flight (results live in two alternating pinned buffers).
    """

    class Handle:
        def __init__(self, event, det_h, ndet_h):
            self.event, self.det_h, self.ndet_h = event, det_h, ndet_h

        def result(self):
            self.event.synchronize()
            out = []
            for b, n in enumerate(self.ndet_h.tolist()):
                d = self.det_h[b, :n]
                out.append(dict(boxes=d[:, :4].clone(), scores=d[:, 4].clone(), labels=d[:, 5].long()))
            return out

    def __init__(self, model):
        self.model = model
        self.engine = model.engine
        self._host = {}
        self._k = 0

    def submit(self, data, filtering=True) -> "PipelinedDetector.Handle":
        m, eng = self.model, self.engine
        if m.backbone.use_image:
            raise NotImplementedError("PipelinedDetector drives the events-only model; use model(data) with images")
        prev = eng.overlap
        eng.overlap = True
        try:
            dec = m.forward_decoded(data, reset=True)
            det, ndet = eng.postprocess(dec, m.conf_threshold, m.nms_threshold, m.width, m.height, filtering=filtering)
            k = self._k & 1
            self._k += 1
            key = (k, tuple(det.shape))
            if key not in self._host:
                self._host[key] = (torch.empty(det.shape, dtype=det.dtype).pin_memory(),
                                   torch.empty(ndet.shape, dtype=ndet.dtype).pin_memory())
            det_h, ndet_h = self._host[key]
            with eng.result_stream():
                det_h.copy_(det, non_blocking=True)
                ndet_h.copy_(ndet, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            eng.fence()                                                  # the slot is free once its results left the device
        finally:
            eng.overlap = prev
        return PipelinedDetector.Handle(ev, det_h, ndet_h)
