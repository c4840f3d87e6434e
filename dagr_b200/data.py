"""Input tensor contract of the hot path + synthetic DSEC-shaped event streams.

`EventBatch` is a duck-typed stand-in for the torch_geometric `Batch` the reference passes to
`DAGR.forward` (built at src/dagr/data/utils.py:6-20, collated with follow_batch=['bbox','bbox0'],
scripts/run_test.py:48): any object exposing the same attributes works.

    raw  (after the DataLoader) : x int8[N,1], pos int16[N,2], t int32[N], batch int64[N],
                                  width/height/time_window int64[B], image u8[B,3,H,W] (optional)
    formatted (format_data)     : x fp32[N,1], pos fp32[N,3] = (x/W, y/H, t/T), t = None,
                                  image fp32 in [0,1]            (src/dagr/utils/buffers.py:33-44)
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np
import torch


class EventBatch:
    _TENSOR_KEYS = ("x", "pos", "t", "batch", "width", "height", "time_window", "image", "bbox", "bbox_batch",
                    "bbox0", "bbox0_batch", "pos_denorm", "t0", "t1")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        if not hasattr(self, "num_graphs") and hasattr(self, "width") and torch.is_tensor(self.width):
            self.num_graphs = int(self.width.numel())

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def _apply(self, fn):
        out = EventBatch()
        for k, v in self.__dict__.items():
            setattr(out, k, fn(v) if torch.is_tensor(v) else copy.copy(v))
        return out

    def clone(self):
        return self._apply(lambda t: t.clone())

    def to(self, device, non_blocking=False):
        return self._apply(lambda t: t.to(device, non_blocking=non_blocking))

    def cuda(self, non_blocking=False):
        return self.to("cuda", non_blocking=non_blocking)

    def cpu(self):
        return self.to("cpu")

    def pin_memory(self):
        return self._apply(lambda t: t.pin_memory())

    def __contains__(self, k):
        return hasattr(self, k)


def format_data(data, normalizer=None):
    """src/dagr/utils/buffers.py:33-44, same semantics (mutates and returns `data`)."""
    if normalizer is None:
        normalizer = torch.stack([data.width[0], data.height[0], data.time_window[0]], dim=-1)
    if hasattr(data, "image") and data.image is not None:
        data.image = data.image.float() / 255.0
    data.pos = torch.cat([data.pos, data.t.view((-1, 1))], dim=-1)
    data.t = None
    data.x = data.x.float()
    data.pos = data.pos / normalizer
    return data


# ------------------------------------------------------------------------------------------------
# synthetic DSEC-shaped streams: deterministic, CPU generator
# ------------------------------------------------------------------------------------------------
def synth_sample(n_events: int, width: int, height: int, seed: int, kind: str = "uniform",
                 time_window: int = 1_000_000, window_us: int = 50_000):
    """one sample: x int16[n], y int16[n], t int32[n] sorted, p int8[n] in {-1,+1}."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    n = int(n_events)
    t0 = time_window - window_us
    t = torch.sort(torch.randint(t0, time_window, (n,), generator=g, dtype=torch.int64)).values.int()
    if kind == "uniform":
        x = torch.randint(0, width, (n,), generator=g)
        y = torch.randint(0, height, (n,), generator=g)
    elif kind == "clustered":
        n_seg = 24
        n_cl = int(0.7 * n)
        seg = torch.randint(0, n_seg, (n_cl,), generator=g)
        cx = torch.rand(n_seg, generator=g) * width
        cy = torch.rand(n_seg, generator=g) * height
        ang = torch.rand(n_seg, generator=g) * np.pi
        length = 40 + torch.rand(n_seg, generator=g) * 160
        speed = (0.2 + torch.rand(n_seg, generator=g) * 1.8) / 1000.0          # px / us
        vang = torch.rand(n_seg, generator=g) * 2 * np.pi
        s = (torch.rand(n_cl, generator=g) - 0.5) * length[seg]
        perp = torch.randn(n_cl, generator=g) * 1.5
        tt = (t[:n_cl].float() - t0)
        # assign the clustered events to random time positions: shuffle which events are clustered
        sel = torch.randperm(n, generator=g)
        is_cl = torch.zeros(n, dtype=torch.bool)
        is_cl[sel[:n_cl]] = True
        tt = (t[is_cl].float() - t0)
        px = cx[seg] + torch.cos(ang[seg]) * s - torch.sin(ang[seg]) * perp + torch.cos(vang[seg]) * speed[seg] * tt
        py = cy[seg] + torch.sin(ang[seg]) * s + torch.cos(ang[seg]) * perp + torch.sin(vang[seg]) * speed[seg] * tt
        x = torch.randint(0, width, (n,), generator=g)
        y = torch.randint(0, height, (n,), generator=g)
        x[is_cl] = px.round().long().remainder(width)
        y[is_cl] = py.round().long().remainder(height)
    else:
        raise ValueError(kind)
    p = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).to(torch.int8)
    return x.to(torch.int16), y.to(torch.int16), t, p


def synth_batch(batch_size: int, n_events: int, width: int = 640, height: int = 480, seed: int = 42,
                kind: str = "uniform", time_window: int = 1_000_000, window_us: int = 50_000,
                with_image: bool = False, ragged: bool = False) -> EventBatch:
    """raw (pre-format_data) batch in the reference's dataset dtypes."""
    xs, ts, ps, bs = [], [], [], []
    g = torch.Generator(device="cpu").manual_seed(int(seed) + 7919)
    for b in range(batch_size):
        n = n_events
        if ragged:
            n = int(n_events * (0.25 + 0.75 * float(torch.rand(1, generator=g))))
        x, y, t, p = synth_sample(n, width, height, seed + b, kind, time_window, window_us)
        xs.append(torch.stack([x, y], dim=1)); ts.append(t); ps.append(p.view(-1, 1))
        bs.append(torch.full((n,), b, dtype=torch.int64))
    kw = dict(x=torch.cat(ps), pos=torch.cat(xs), t=torch.cat(ts), batch=torch.cat(bs),
              width=torch.full((batch_size,), width, dtype=torch.int64),
              height=torch.full((batch_size,), height, dtype=torch.int64),
              time_window=torch.full((batch_size,), time_window, dtype=torch.int64),
              num_graphs=batch_size)
    if with_image:
        kw["image"] = torch.randint(0, 256, (batch_size, 3, height, width), generator=g, dtype=torch.uint8)
    kw["dims"] = (int(width), int(height), int(time_window))      # host copy of width/height/time_window: no device read per forward
    return EventBatch(**kw)
