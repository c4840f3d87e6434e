#!/usr/bin/env python
"""tiny driver for ncu captures: N forwards of the config-2 workload (no timing, no baseline)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from dagr_b200.data import format_data, synth_batch
from dagr_b200.utils.args import default_args
from dagr_b200.model.dagr import DAGR
from tests.helpers import randomize_bn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
B, EV = 8, 300000
torch.manual_seed(0)
m = randomize_bn(DAGR(default_args("s", batch_size=B), height=480, width=640).eval()).cuda()
d = format_data(synth_batch(B, EV, 640, 480, seed=2042, kind=kind).cuda())
for i in range(n):
    dec = m.forward_decoded(d)
    det, nd = m.engine.postprocess(dec, 0.001, 0.65, 640, 480)
torch.cuda.synchronize()
print("ok", nd.tolist())
