#!/usr/bin/env python
"""per-op CUDA-event profile of the config-3 shape (dagr-s + image fusion, 8 x 300k events, 640x480)."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from dagr_b200.data import format_data, synth_batch
from dagr_b200.utils.args import default_args
from dagr_b200.model.dagr import DAGR
from tests.helpers import randomize_bn

B, EV = 8, int(sys.argv[1]) if len(sys.argv) > 1 else 300000
net = sys.argv[2] if len(sys.argv) > 2 else "resnet50"
torch.manual_seed(0)
args = default_args("s", batch_size=B, use_image=True, img_net=net)
m = randomize_bn(DAGR(args, height=480, width=640).eval()).cuda()
d = format_data(synth_batch(B, EV, 640, 480, seed=2042, with_image=True).cuda())
for i in range(3):
    m(d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    m(d)
torch.cuda.synchronize()
print("ms/forward", (time.perf_counter() - t0) / 5 * 1e3)
# trunk alone
with torch.no_grad():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        feats, outs = m.backbone.net(d.image.float())
    e1.record()
    torch.cuda.synchronize()
print("trunk ms", e0.elapsed_time(e1) / 5)
m.engine.prof = {}
for i in range(5):
    m(d)
torch.cuda.synchronize()
for k, v in sorted(m.engine.prof_summary().items(), key=lambda kv: -kv[1]["ms"] * kv[1]["calls"])[:16]:
    print(f"{k:28s} {v['ms']:.4f} ms x {v['calls'] / 5:.0f}")
