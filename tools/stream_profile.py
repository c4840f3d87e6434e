#!/usr/bin/env python
"""per-op timing of incremental steps (config-5 shape): 1 stream, 50 ms history, 1 ms chunks."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from dagr_b200.asynchronous import AsyncDAGR
from dagr_b200.data import EventBatch, format_data, synth_batch
from dagr_b200.model.dagr import DAGR
from dagr_b200.utils.args import default_args
from tests.helpers import randomize_bn

size = sys.argv[1] if len(sys.argv) > 1 else "l"
W, H, T = 640, 480, 1_000_000
torch.manual_seed(0)
m = randomize_bn(DAGR(default_args(size, batch_size=1), height=H, width=W).eval()).cuda()
d = format_data(synth_batch(1, 100_000, W, H, seed=99, kind="uniform", window_us=100_000).cuda())
t_us = (d.pos[:, 2].double() * T).round(); t0 = float(t_us.min())
eng = AsyncDAGR(m)
def chunk(lo, hi):
    c = (t_us >= t0 + lo) & (t_us < t0 + hi)
    return EventBatch(x=d.x[c], pos=d.pos[c], batch=d.batch[c], width=d.width, height=d.height, time_window=d.time_window, num_graphs=1)
eng.step(chunk(0, 50_000))
for k in range(6):
    eng.step(chunk(50_000 + 1000 * k, 51_000 + 1000 * k))
torch.cuda.synchronize()
# host time vs device time of 10 steps
chs = [chunk(56_000 + 1000 * k, 57_000 + 1000 * k) for k in range(10)]
torch.cuda.synchronize(); t1 = time.perf_counter()
for c in chs: eng.step(c)
torch.cuda.synchronize(); t2 = time.perf_counter()
print("wall per step (ms):", (t2 - t1) / 10 * 1e3)
chs = [chunk(66_000 + 1000 * k, 67_000 + 1000 * k) for k in range(10)]
t1 = time.perf_counter()
for c in chs: eng.step_decoded(c, batch_size=1)
t_host = time.perf_counter() - t1
torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue per step_decoded (ms):", t_host / 10 * 1e3, " incl. drain:", (t2 - t1) / 10 * 1e3)
m.engine.prof = {}
for k in range(5): eng.step(chunk(76_000 + 1000 * k, 77_000 + 1000 * k))
torch.cuda.synchronize()
ps = m.engine.prof_summary(); m.engine.prof = None
tot = sum(v["ms"] * v["calls"] / 5 for v in ps.values())
print("sum of per-op device ms per step (eager, with events):", tot)
for k, v in sorted(ps.items(), key=lambda kv: -kv[1]["ms"] * kv[1]["calls"])[:14]:
    print(f"  {k:28s} {v['ms']:.4f} ms x {v['calls'] / 5:.0f}")
