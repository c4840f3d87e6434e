#!/usr/bin/env python
"""Generate tests/golden/downsample_golden.npz from the REFERENCE's own down-sampler (build container only).

/root/reference/scripts/downsample_events.py imports h5py / hdf5plugin / dsec_det at module level (absent here) for its
file I/O; they are stubbed so that the numba function `downsample_events` itself runs unmodified.  Three chunks are
pushed through with the change map carried over, as the script's main loop does (:146-153)."""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

for name in ("hdf5plugin", "h5py", "tqdm", "dsec_det", "dsec_det.io"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["h5py"].File = object                                    # only used in a type annotation
sys.modules["dsec_det.io"].extract_from_h5_by_index = None
sys.modules["dsec_det.io"].get_num_events = None
spec = importlib.util.spec_from_file_location("ref_downsample", "/root/reference/scripts/downsample_events.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(7)
out = {}
for case, (iw, ih, ow, oh, n) in enumerate([(640, 480, 320, 240, 6000), (64, 48, 32, 24, 5000), (96, 48, 32, 24, 4000)]):
    cm = None
    for chunk in range(3):
        ev = dict(x=rng.integers(0, iw, n).astype("uint16"), y=rng.integers(0, ih, n).astype("uint16"),
                  p=(2 * rng.integers(0, 2, n) - 1).astype("int8"), t=np.sort(rng.integers(0, 50000, n)).astype("int64") + 50000 * chunk)
        if case == 1 and chunk == 1:                                    # hot pixels: long runs on a few cells
            ev["x"][: n // 2] = 10
            ev["y"][: n // 2] = 7
            ev["p"][: n // 2 : 3] = 1
        res, cm = ref.downsample_events({k: v.copy() for k, v in ev.items()}, ih, iw, oh, ow, change_map=cm)
        for k, v in ev.items():
            out[f"c{case}_k{chunk}_in_{k}"] = v
        for k, v in res.items():
            out[f"c{case}_k{chunk}_out_{k}"] = v
        out[f"c{case}_k{chunk}_map"] = cm.copy()
    out[f"c{case}_shape"] = np.array([iw, ih, ow, oh])
np.savez_compressed(Path(__file__).parent / "downsample_golden.npz", **out)
print("wrote", len(out), "arrays")
