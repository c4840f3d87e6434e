"""The oracle (oracle/ref_model.py, oracle/ref_ops.py) against fixtures produced by the REFERENCE's own code:
tests/golden/reference_forward_*.npz hold the outputs of /root/reference/src/dagr's unmodified DAGR.forward
(net.py, dagr.py, pooling.py, conv.py, components.py, spline_conv.py incl. init_lut / message_lut / to_dense, ev_tgn.py,
model/utils.py incl. postprocess + the real torchvision NMS), run in the build container with the stand-ins of
tests/golden/ref_shim.py for the absent third-party packages.  tests/golden/reference_async_helpers.npz holds outputs of
the pure-torch helpers of src/dagr/asynchronous (cartesian.__edge_attr, max_pool.pool_edge / compute_attrs /
__get_global_cluster_index, base/utils._to_hom / _from_hom).  No file under /root/reference is read here."""
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.helpers import REFERENCE_FIXTURES, assert_close, image_branch_cpu, load_reference_fixture

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", REFERENCE_FIXTURES)
def test_oracle_equals_the_references_own_forward(name):
    from oracle.ref_model import RefModel
    model, args, data, exp, meta = load_reference_fixture(name)          # also pins the state_dict key layout (checksum)
    kw = {}
    if meta["use_image"]:
        feats, outs = image_branch_cpu(model, data)
        kw = dict(image_feats=feats, image_outs=outs)
    o = RefModel(model.state_dict(), args, meta["H"], meta["W"]).forward(data.x, data.pos, data.batch, meta["B"], **kw)
    assert torch.equal(o["edge_index"], exp["edge_index"])
    tol = 1e-5 if args.num_scales == 2 else 1e-4                         # single-scale fixture: basis form upstream, LUT form here
    assert_close(o["x1a"], exp["x1a"], tol=tol, what="golden x1a")
    assert_close(o["x1"], exp["x1"], tol=tol, what="golden x1")
    for lv in range(4):
        a, b = o["levels"][lv], exp["levels"][lv]
        assert a["x"].shape == b["x"].shape and torch.equal(a["batch"], b["batch"]), f"level {lv}"
        keep = ~b["ambiguous"]
        assert torch.equal(a["pos"][keep, :2], b["pos"][keep, :2]), f"level {lv}: pooled (rounded) positions"
        assert_close(a["pos"][:, 2], b["pos"][:, 2], tol=1e-6, what="golden pooled mean t")      # fp64 mean here, fp32 upstream
        assert torch.equal(a["edge_index"], b["edge_index"]), f"level {lv}: coarse edges"
        assert_close(a["x"], b["x"], tol=tol, what=f"golden level {lv} features")
    assert_close(o["out3"], exp["out3"], tol=tol, what="golden out3")
    assert_close(o["out4"], exp["out4"], tol=tol, what="golden out4")
    if not meta["use_image"]:                                            # upstream hooks see the maps before the CNN maps are added
        for k in range(args.num_scales):
            for nm in ("cls", "reg", "obj"):
                assert_close(o["dense"][k][nm], exp["dense"][k][nm], tol=tol, what=f"golden dense {nm}{k + 1}")
    assert_close(o["decoded"], exp["decoded"], tol=tol, what="golden decoded")
    for b in range(meta["B"]):
        assert len(o["detections"][b]["boxes"]) == len(exp["detections"][b]["boxes"])
        assert_close(o["detections"][b]["boxes"], exp["detections"][b]["boxes"], tol=tol, what="golden boxes")
        assert torch.equal(o["detections"][b]["labels"], exp["detections"][b]["labels"])


def test_oracle_equals_the_references_async_helpers():
    """src/dagr/asynchronous/cartesian.py:6-16, max_pool.py:13-26,245-252, base/utils.py:23-31 (pure torch upstream)."""
    from oracle import ref_ops as R
    g = np.load(GOLD / "reference_async_helpers.npz")
    t = lambda k: torch.from_numpy(g[k])
    pos, ei = t("pos"), t("edge_index").long()
    for tag, mx in (("a", float(g["max_a"])), ("b", float(g["max_b"]))):
        assert torch.equal(R.cartesian(pos, ei, mx), t(f"edge_attr_{tag}"))
    # pool_edge == the coarse-edge step of Pooling.forward (pooling.py:58-64)
    cl = t("cluster").long()
    e = cl[ei]
    e = e[:, e[0] != e[1]].unique(dim=-1)
    assert torch.equal(e, t("pooled_edges").long())
    assert torch.equal(R.cartesian(t("cpos"), t("pooled_edges").long(), float(g["max_b"])), t("pooled_attr"))
    # fixed-slot cell index of the asynchronous pooling == grid_cluster on (x, y) of one sample
    vs = t("voxel_size")
    nx = int((1 / vs[0] + 1e-3).long())
    c = R.grid_cluster(torch.cat([pos[:, :2], torch.zeros(len(pos), 2)], 1), torch.cat([vs[:2], torch.ones(2)]),
                       torch.zeros(4), torch.tensor([0.9999999, 0.9999999, 0.9999999, 0.0]))
    ref_c = t("global_cluster").long()
    assert torch.equal(c % nx + nx * (c // nx), ref_c) or torch.equal(c, ref_c)
    # homogeneous mean (the asynchronous pooling's position update) == scatter_mean up to the 1e-9 guard
    hom = t("hom")
    assert torch.equal(hom[:, :-1], pos) and bool((hom[:, -1] == 1).all())
    assert_close(t("from_hom"), R.scatter_mean(pos, cl, int(cl.max()) + 1), tol=1e-6, what="from_hom vs scatter_mean")
