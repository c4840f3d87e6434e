"""Does the captured coarse stack survive `ncu` in its default per-node graph profiling mode?
usage: ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 python tools/ncu_graph_probe.py <overlap 0|1> <scales 1|2>
Runs 5 forwards of dagr-s (2 eager, capture, 2 replays) on a small batch and prints `probe ok`."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import make_model, make_inputs

overlap, scales = (int(v) for v in sys.argv[1:3])
model, args = make_model("s", 480, 640, batch_size=2, num_scales=scales)
model.cuda()
eng = model.engine
eng.overlap = bool(overlap)
raw, data = make_inputs(2, 30000, 640, 480, seed=5)
data = data.to("cuda") if hasattr(data, "to") else data
for i in range(5):
    out = model(data)
    torch.cuda.synchronize()
print("probe ok", overlap, scales, len(out[0]) if isinstance(out, (list, tuple)) else type(out))
