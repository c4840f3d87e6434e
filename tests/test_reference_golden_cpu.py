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


# ------------------------------------------------------------------------------------------------------------------
# The arithmetic inside the absent third-party packages (torch_spline_conv, torch_cluster, torch_scatter) has no fixture
# from the packages themselves; the oracle's restatements are cross-checked against INDEPENDENT implementations of the same
# published definitions (scipy B-splines, numpy / pandas group-bys) so that they are not only self-consistent.
# ------------------------------------------------------------------------------------------------------------------
def test_spline_basis_matches_scipy_bspline_design_matrix():
    """degree-1 open B-spline basis on kernel_size uniform knots == scipy.interpolate.BSpline.design_matrix (k = 1)."""
    from scipy.interpolate import BSpline
    from oracle import ref_ops as R
    ks = 5
    g = torch.Generator().manual_seed(2)
    pseudo = torch.rand(4000, 2, generator=g, dtype=torch.float64)
    pseudo[:8] = torch.tensor([[0.0, 0.0], [0.25, 0.5], [0.5, 0.5], [0.999999, 0.3], [0.125, 0.875], [0.75, 0.0], [0.3, 0.7], [0.6, 0.1]])
    basis, index = R.spline_basis(pseudo, ks, True, 1)
    dense = torch.zeros(len(pseudo), ks * ks, dtype=torch.float64)
    dense.scatter_add_(1, index, basis)
    # open spline, degree 1: ks hat functions on the uniform knots 0, 1/(ks-1), ..., 1 (clamped ends)
    t = np.concatenate([[0.0], np.linspace(0.0, 1.0, ks), [1.0]])
    bx = BSpline.design_matrix(pseudo[:, 0].numpy(), t, 1).toarray()          # [E, ks]
    by = BSpline.design_matrix(pseudo[:, 1].numpy(), t, 1).toarray()
    want = (by[:, :, None] * bx[:, None, :]).reshape(len(pseudo), ks * ks)    # slot = kx + ks * ky
    assert np.abs(dense.numpy() - want).max() < 1e-12
    assert np.allclose(basis.sum(1).numpy(), 1.0)


def test_grid_cluster_and_scatter_match_numpy_and_pandas():
    import pandas as pd
    from oracle import ref_ops as R
    g = torch.Generator().manual_seed(4)
    n, B = 5000, 3
    pos = torch.rand(n, 3, generator=g)
    batch = torch.randint(0, B, (n,), generator=g)
    size = torch.tensor([1.0 / 56, 1.0 / 40, 1.0, 1.0])
    pos4 = torch.cat([pos, batch.float().view(-1, 1)], 1)
    c = R.grid_cluster(pos4, size, torch.zeros(4), torch.tensor([0.9999999, 0.9999999, 0.9999999, B - 1.0]))
    # voxel id = x-cell + 56 * (y-cell + 40 * (t-cell + 1 * batch)), cells by fp32 division and truncation
    cell = np.trunc(pos4.numpy() / size.numpy()).astype(np.int64)
    want = cell[:, 0] + 56 * (cell[:, 1] + 40 * (cell[:, 2] + 1 * cell[:, 3]))
    assert np.array_equal(c.numpy(), want)
    uniq, cl, _, _ = R.consecutive_cluster(c)
    x = torch.randn(n, 7, generator=g)
    df = pd.DataFrame(x.numpy()).assign(cl=cl.numpy())
    assert np.array_equal(R.scatter_max(x, cl, len(uniq)).numpy(), df.groupby("cl").max().to_numpy().astype(np.float32))
    mean = df.groupby("cl").mean().to_numpy()
    assert np.abs(R.scatter_mean(x, cl, len(uniq)).numpy() - mean).max() < 1e-5
