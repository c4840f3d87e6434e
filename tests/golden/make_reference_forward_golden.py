#!/usr/bin/env python
"""Generate tests/golden/reference_forward_*.npz by running the REFERENCE's own DAGR.forward (build container only).

The reference's model files are imported unmodified from /root/reference/src; the third-party packages they need are
the stand-ins of tests/golden/ref_shim.py (see its header for what that pins and what it does not).  Weights come from
tests.helpers.golden_weights (deterministic rule shared with the tests), so a fixture holds inputs + outputs only:

  edge_index, conv_block1.conv_block1 / conv_block1 outputs, the four pooled graphs (x, pos, batch, edge_index),
  layer4 / layer5 outputs, the six dense head maps, decoded outputs [B,175,5+nc] and the post-processed detections
  (the reference's postprocess_network_output with the real torchvision NMS).

    python tests/golden/make_reference_forward_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from tests.golden import ref_shim                                        # noqa: E402

ref_shim.install()
from dagr.model.networks import net as ref_net                            # noqa: E402  (the reference's files)
from dagr.model.networks.dagr import DAGR as RefDAGR                      # noqa: E402
from torch_geometric.data import Batch                                    # noqa: E402  (stand-in)
import torchvision                                                        # noqa: E402

from dagr_b200.data import format_data, synth_batch                       # noqa: E402
from dagr_b200.utils.args import default_args                             # noqa: E402
from tests.helpers import golden_weights                                  # noqa: E402

# `img_net(pretrained=True)` (net.py:44) would download weights: same architectures, no download
for _n in ("resnet18", "resnet34", "resnet50"):
    setattr(ref_net, _n, (lambda f: (lambda pretrained=True: f(weights=None)))(getattr(torchvision.models, _n)))

CASES = [
    # name, size, dataset, W, H, B, events/sample, kind, use_image, seed
    ("events_s_320x215", "s", "dsec", 320, 215, 2, 9000, "clustered", False, 101),
    ("events_n_ncaltech_240x180", "n", "ncaltech101", 240, 180, 1, 10000, "uniform", False, 102),
    ("image_n_240x180", "n", "dsec", 240, 180, 2, 6000, "clustered", True, 103),
]


def run(name, size, dataset, W, H, B, n, kind, use_image, seed):
    args = default_args(size, dataset=dataset, batch_size=B, use_image=use_image, img_net="resnet18")
    torch.manual_seed(0)
    model = RefDAGR(args, height=H, width=W).eval()
    checksum = golden_weights(model, seed)
    if args.num_scales == 2:
        model.cache_luts(width=W, height=H, radius=args.radius)        # run_test.py:59
    # num_scales == 1 (N-Caltech101): cache_luts would give the single head scale the LUT geometry of pool3 although it
    # runs on the pool4 graph (quirk Q6, dagr.py:51-72); that configuration is evaluated in the spline-basis form instead
    raw = synth_batch(B, n, W, H, seed=seed, kind=kind, with_image=use_image, ragged=True)
    d = format_data(raw.clone())
    data = Batch(x=d.x.clone(), pos=d.pos.clone(), batch=d.batch.clone(), width=d.width, height=d.height,
                 time_window=d.time_window, num_graphs=B)
    if use_image:
        data.image = d.image.clone()
    cap = {}

    def grab(key, fn):
        def hook(mod, inp, out):
            cap[key] = fn(out)
        return hook

    bb, hd = model.backbone, model.head
    hs = [bb.events_to_graph.register_forward_hook(grab("edge_index", lambda o: o.edge_index.clone())),
          bb.conv_block1.conv_block1.register_forward_hook(grab("x1a", lambda o: o.x.clone())),
          bb.conv_block1.register_forward_hook(grab("x1", lambda o: o.x.clone())),
          bb.layer4.register_forward_hook(grab("out3", lambda o: o.x.clone())),
          bb.layer5.register_forward_hook(grab("out4", lambda o: o.x.clone())),
          hd.register_forward_hook(grab("decoded", lambda o: o.clone()))]
    for i in range(4):
        hs.append(getattr(bb, f"pool{i + 1}").register_forward_hook(
            grab(f"level{i}", lambda o: dict(x=o.x.clone(), pos=o.pos.clone(), batch=o.batch.clone(), edge_index=o.edge_index.clone()))))
    for k in range(args.num_scales):
        for nm in ("cls", "reg", "obj"):
            hs.append(getattr(hd, f"{nm}_pred{k + 1}").register_forward_hook(grab(f"dense_{nm}{k + 1}", lambda o: o.clone())))
    if use_image:
        hs.append(bb.net.register_forward_hook(grab("image", lambda o: ([f.clone() for f in o[0]], [f.clone() for f in o[1]]))))
        hs.append(hd.cnn_head.register_forward_hook(grab("cnn", lambda o: {k: [t.clone() for t in v] for k, v in o.items()})))
    with torch.no_grad():
        dets = model(data)[0]
    for h in hs:
        h.remove()
    out = dict(meta=np.array([W, H, B, seed, int(use_image)]), size=np.array(size), dataset=np.array(dataset),
               weight_checksum=np.array(checksum), x=d.x.numpy(), pos=d.pos.numpy(), batch=d.batch.numpy().astype(np.int32))
    if use_image:
        out["image_u8"] = raw.image.numpy()
    out["edge_index"] = cap["edge_index"].numpy().astype(np.int32)
    for k in ("x1a", "x1", "out3", "out4", "decoded"):
        out[k] = cap[k].numpy()
    for i in range(4):
        lv = cap[f"level{i}"]
        out[f"level{i}_x"] = lv["x"].numpy()
        out[f"level{i}_pos"] = lv["pos"].numpy()
        out[f"level{i}_batch"] = lv["batch"].numpy().astype(np.int32)
        out[f"level{i}_edge_index"] = lv["edge_index"].numpy().astype(np.int32)
    # the dense head maps: the hooks saw them BEFORE the CNN maps are added in place (dagr.py:219-222)
    for k in range(args.num_scales):
        for nm in ("cls", "reg", "obj"):
            out[f"dense_{nm}{k + 1}"] = cap[f"dense_{nm}{k + 1}"].numpy()
    for b, det in enumerate(dets):
        out[f"det{b}_boxes"] = det["boxes"].numpy()
        out[f"det{b}_scores"] = det["scores"].numpy()
        out[f"det{b}_labels"] = det["labels"].numpy().astype(np.int32)
    # the oracle on the same inputs: must agree with the reference's own code here, and supplies the flags of the voxels
    # whose mean position lies within float noise of a pixel boundary (the reference's fp32 mean is order dependent there)
    from oracle.ref_model import RefModel
    from tests.helpers import rel_err
    kw = {}
    if use_image:
        kw = dict(image_feats=cap["image"][0], image_outs=cap["cnn"])
    o = RefModel(model.state_dict(), args, H, W).forward(d.x, d.pos, d.batch, B, **kw)
    assert torch.equal(o["edge_index"], cap["edge_index"])
    errs = dict(x1=rel_err(o["x1"], cap["x1"]), decoded=rel_err(o["decoded"], cap["decoded"]))
    for i in range(4):
        out[f"level{i}_ambiguous"] = o["levels"][i]["ambiguous"].numpy()
        errs[f"level{i}"] = rel_err(o["levels"][i]["x"], cap[f"level{i}"]["x"]) if o["levels"][i]["x"].shape == cap[f"level{i}"]["x"].shape else -1
    print("   oracle vs reference:", {k: f"{v:.1e}" for k, v in errs.items()})
    np.savez_compressed(HERE / f"reference_forward_{name}.npz", **out)
    nd = [len(x["boxes"]) for x in dets]
    print(f"{name}: N={len(d.x)} E={out['edge_index'].shape[1]} levels={[out[f'level{i}_x'].shape[0] for i in range(4)]} detections={nd} "
          f"checksum={checksum:.6e}")


if __name__ == "__main__":
    for c in CASES:
        run(*c)
