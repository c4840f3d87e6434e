#!/usr/bin/env python
"""kernel timeline of a few config-2 steps (torch.profiler / CUPTI): start, duration, stream of every kernel, so that
gaps (host-bound phases) and cross-stream overlap are visible.  usage: python tools/timeline.py [overlap 0|1] [steps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from torch.profiler import profile, ProfilerActivity
from dagr_b200.data import format_data, synth_batch
from dagr_b200.utils.args import default_args
from dagr_b200.model.dagr import DAGR
from tests.helpers import randomize_bn

overlap = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, EV = 8, 300000
torch.manual_seed(0)
m = randomize_bn(DAGR(default_args("s", batch_size=B), height=480, width=640).eval()).cuda()
ds = [format_data(synth_batch(B, EV, 640, 480, seed=2042 + i).cuda()) for i in range(3)]
eng = m.engine
eng.overlap = bool(overlap)


def step(i):
    dec = m.forward_decoded(ds[i % 3])
    return eng.postprocess(dec, 0.001, 0.65, 640, 480)


for i in range(6):
    step(i)
eng.join(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        step(i)
    eng.join(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = t0
for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    if d >= 15 or gap > 20:
        print(f"{s/1000:9.3f} ms  dur {d:8.1f} us  gap {gap:7.1f} us  stream {getattr(e, 'stream', '?')}  {e.name[:70]}")
    prev_end = max(prev_end, e.time_range.end)
print("total", (evs[-1].time_range.end - t0) / 1000 / steps, "ms/step", "kernels", len(evs) / steps)
