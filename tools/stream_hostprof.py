#!/usr/bin/env python
"""cProfile of the host side of incremental steps (config-5 shape, 1 ms chunks)."""
import cProfile, pstats, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from dagr_b200.asynchronous import AsyncDAGR
from dagr_b200.data import EventBatch, format_data, synth_batch
from dagr_b200.model.dagr import DAGR
from dagr_b200.utils.args import default_args
from tests.helpers import randomize_bn

W, H, T = 640, 480, 1_000_000
torch.manual_seed(0)
m = randomize_bn(DAGR(default_args("l", batch_size=1), height=H, width=W).eval()).cuda()
d = format_data(synth_batch(1, 300_000, W, H, seed=99, kind="uniform", window_us=300_000).cuda())
t_us = (d.pos[:, 2].double() * T).round(); t0 = float(t_us.min())
eng = AsyncDAGR(m)
def chunk(lo, hi):
    c = (t_us >= t0 + lo) & (t_us < t0 + hi)
    return EventBatch(x=d.x[c], pos=d.pos[c], batch=d.batch[c], width=d.width, height=d.height, time_window=d.time_window, num_graphs=1,
                      dims=(W, H, T))
eng.step(chunk(0, 50_000))
for k in range(6):
    eng.step(chunk(50_000 + 1000 * k, 51_000 + 1000 * k))
chs = [chunk(56_000 + 1000 * k, 57_000 + 1000 * k) for k in range(200)]
torch.cuda.synchronize()
pr = cProfile.Profile()
t1 = time.perf_counter()
pr.enable()
for c in chs:
    eng.step(c)
pr.disable()
torch.cuda.synchronize()
print("wall per step (ms):", (time.perf_counter() - t1) / len(chs) * 1e3)
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
