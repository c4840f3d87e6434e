"""Writes tests/golden/forward_dagr_n_240x180.pt: a small seeded forward of the ORACLE (oracle/ref_model.py, CPU)
-- inputs and every intermediate the GPU parity test checks.  The reference itself cannot produce vectors
(torch_geometric & co. are not installable offline; no checkpoint or fixtures ship with it), so the
fixture pins the oracle restatement as of this commit: edge_index (bit-exact graph, itself pinned by the
reference's own CUDA kernels), layer-1 activations, pooled levels, dense head maps, decoded outputs, detections.
Weights are NOT stored: they are regenerated from the same seed (tests/helpers.make_model)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import torch
from tests.helpers import make_inputs, make_model
from oracle.ref_model import RefModel

W, H, B, N = 240, 180, 2, 1500
model, args = make_model("n", H, W, seed=0)
raw, data = make_inputs(B, N, W, H, seed=123, kind="clustered")
ref = RefModel(model.state_dict(), args, H, W)
o = ref.forward(data.x, data.pos, data.batch, B)
fix = dict(meta=dict(W=W, H=H, B=B, n=N, size="n", model_seed=0, input_seed=123, kind="clustered"),
           x=data.x, pos=data.pos, batch=data.batch,
           edge_index=o["edge_index"].int(), x1a=o["x1a"].half().float(), x1=o["x1"],
           level_pos=[l["pos"][:, :2] for l in o["levels"]], level_batch=[l["batch"].int() for l in o["levels"]],
           level_x=[l["x"] for l in o["levels"][:2]], level_edges=[l["edge_index"].int() for l in o["levels"]],
           level_ambiguous=[l["ambiguous"] for l in o["levels"]],
           decoded=o["decoded"], n_det=[len(d["boxes"]) for d in o["detections"]],
           weight_checksum=float(sum(v.double().abs().sum() for v in model.state_dict().values() if v.dtype.is_floating_point)))
fix["x1a"] = o["x1a"]
torch.save(fix, Path(__file__).with_name("forward_dagr_n_240x180.pt"))
print("wrote", {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in fix.items()})
