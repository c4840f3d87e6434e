#!/usr/bin/env python
"""Generate tests/golden/records_golden.npz from the REFERENCE's own src/dagr/utils/buffers.py (build container only).
Its sibling module coco_eval (pycocotools) is stubbed; the functions exercised here do not use it."""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

for name, path in (("refdagr", []), ("refdagr.utils", ["/root/reference/src/dagr/utils"])):
    m = types.ModuleType(name)
    m.__path__ = path
    sys.modules[name] = m
stub = types.ModuleType("refdagr.utils.coco_eval")
stub.evaluate_detection = None
sys.modules["refdagr.utils.coco_eval"] = stub
spec = importlib.util.spec_from_file_location("refdagr.utils.buffers", "/root/reference/src/dagr/utils/buffers.py")
ref = importlib.util.module_from_spec(spec)
sys.modules["refdagr.utils.buffers"] = ref
spec.loader.exec_module(ref)

g = torch.Generator().manual_seed(3)
out = {}
dets, gts = [], []
for i, n in enumerate([5, 0, 7, 3]):
    xy = torch.rand(n, 2, generator=g) * 500 - 20
    wh = torch.rand(n, 2, generator=g) * 120
    boxes = torch.cat([xy, xy + wh], dim=1)
    dets.append(dict(boxes=boxes, scores=torch.rand(n, generator=g), labels=torch.randint(0, 2, (n,), generator=g)))
    gts.append(dict(boxes=boxes.flip(0).clone(), labels=torch.randint(0, 2, (n,), generator=g)))
    for k, v in dets[-1].items():
        out[f"det{i}_{k}"] = v.numpy().copy()
    for k, v in gts[-1].items():
        out[f"gt{i}_{k}"] = v.numpy().copy()
seqs, ts = ["zurich_a", "zurich_a", "interlaken_b", "zurich_a"], [1000, 51000, 7, 101000]
buf = ref.DetectionBuffer(height=430, width=640, classes=["car", "pedestrian"])
buf.update([{k: v.clone() for k, v in d.items()} for d in dets], [{k: v.clone() for k, v in d.items()} for d in gts], "dsec")
cd, cg = buf.compile(seqs, ts)
for s, v in cd.items():
    out[f"compiled_det_{s}"] = v
for s, v in cg.items():
    out[f"compiled_gt_{s}"] = v
flt = ref.filter_bboxes([{k: v.clone() for k, v in d.items()} for d in dets], 430, 640)
for i, d in enumerate(flt):
    for k, v in d.items():
        out[f"filt{i}_{k}"] = v.numpy().copy()
db = ref.DictBuffer()
for d in ({"a": 1.0, "b": 4.0}, {"a": 3.0, "b": 0.0}, {"a": 8.0, "b": 2.0}):
    db.update(d)
out["dictbuffer"] = np.array([db.compute()["a"], db.compute()["b"]])
np.savez_compressed(Path(__file__).parent / "records_golden.npz", **out)
print("wrote", len(out), "arrays")
