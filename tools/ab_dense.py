"""A/B of the dense-voxel work lists (Engine.dense_worklists) on the headline workload: per-op CUDA-event times."""
import sys, json, torch
sys.path.insert(0, ".")
from dagr_b200.data import format_data, synth_batch
from dagr_b200.model.dagr import DAGR
from dagr_b200.utils.args import default_args
from tests.helpers import randomize_bn
kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
torch.manual_seed(0)
m = randomize_bn(DAGR(default_args("s", batch_size=8), height=480, width=640).eval()).cuda()
d = format_data(synth_batch(8, 300000, 640, 480, seed=2042, kind=kind).to("cuda"))
out = {}
for flag in (True, False, True, False):
    m.engine.dense_worklists = flag
    for _ in range(3):
        m.forward_decoded(d)
    m.engine.prof = {}
    for _ in range(10):
        m.forward_decoded(d)
    torch.cuda.synchronize()
    p = m.engine.prof_summary(); m.engine.prof = None
    out.setdefault(str(flag), []).append({k: round(p[k]["ms"], 4) for k in ("l1_build", "l1_conv_b_pool_voxel", "graph_sort")})
print(json.dumps(dict(kind=kind, per_op_ms=out)))
