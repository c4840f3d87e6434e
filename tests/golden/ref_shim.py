"""TEST INFRASTRUCTURE (build container only): lets the REFERENCE's own model code run here.

`install()` puts minimal stand-ins for the third-party packages the reference imports at module level
(torch_geometric, torch_scatter, torch_cluster, torch_spline_conv, yolox, and the two CUDA extensions) into
`sys.modules` and adds /root/reference/src to `sys.path`, so that `from dagr.model.networks.dagr import DAGR`
imports the reference's UNMODIFIED files (net.py, dagr.py, pooling.py, conv.py, components.py, spline_conv.py,
ev_tgn.py, ev_graph.py, graph/utils.py, model/utils.py, asynchronous/*.py).  Everything the reference repo itself
implements -- orchestration, LUT construction and lookup, to_dense, edge pooling, rounding, head, decode, post-process
(with the real torchvision NMS) -- therefore runs as written upstream and pins the oracle / the CUDA path.

What the stand-ins restate (and what therefore stays UNPINNED, stated in DESIGN.md): the arithmetic INSIDE the absent
packages -- torch_spline_conv.spline_basis, torch_cluster.grid_cluster, torch_scatter.scatter_*, PyG's
SplineConv.propagate / ToSparseTensor / Cartesian / BatchNorm plumbing -- taken from oracle/ref_ops.py, and the
radius-graph CUDA extension, replaced by oracle/graph_oracle.c (itself pinned against the reference's ev_graph.cu on
the GPU box, tests/test_gpu_parity.py::test_graph_vs_reference_cuda_kernels).
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF_SRC = "/root/reference/src"


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_mod(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


# ------------------------------------------------------------------------------------------------------------------
# torch_geometric.data
# ------------------------------------------------------------------------------------------------------------------
class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __getitem__(self, k):
        return getattr(self, k, None)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    def __iter__(self):
        return iter(list(self.__dict__.items()))

    def __contains__(self, k):
        return hasattr(self, k)

    @property
    def num_nodes(self):
        for k in ("x", "pos"):
            v = getattr(self, k, None)
            if v is not None:
                return v.shape[0]
        return 0

    @property
    def num_edges(self):
        return self.edge_index.shape[1]

    def clone(self):
        return type(self)(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})

    def to(self, *a, **kw):
        return self


class Batch(Data):
    @classmethod
    def from_data_list(cls, lst):
        d = lst[0]
        out = cls(**d.__dict__)
        out.batch = torch.zeros(d.num_nodes, dtype=torch.long)
        return out


# ------------------------------------------------------------------------------------------------------------------
# PyG transforms / layers (plumbing restated from the published sources; arithmetic from oracle/ref_ops.py)
# ------------------------------------------------------------------------------------------------------------------
class _Adj:
    """what SplineConv.propagate needs from torch_sparse.SparseTensor: (dst, src, value) sorted by (dst, src)."""

    def __init__(self, dst, src, value, n):
        self.dst, self.src, self.value, self.n = dst, src, value, n

    def numel(self):
        return self.dst.numel()


class ToSparseTensor:
    """torch_geometric 2.0.x semantics (the `attr=` keyword the reference passes, spline_conv.py:12, exists from 2.0 on):
    edge_index AND every edge-level attribute are sorted consistently by (dst, src) inside the data object (sort_edge_index,
    sort_by_row=False), then adj_t is built from the sorted arrays.  (The 1.x transform permuted the attributes but left
    edge_index alone; with that behaviour later shallow copies would pair attributes with the wrong edges.)"""

    def __init__(self, attr="edge_weight", remove_edge_index=True, fill_cache=True):
        self.attr, self.remove_edge_index = attr, remove_edge_index

    def __call__(self, data):
        (row, col), N, E = data.edge_index, data.num_nodes, data.num_edges
        perm = torch.argsort(col * N + row, stable=True)
        for key, item in data:
            if key != "edge_index" and torch.is_tensor(item) and item.dim() > 0 and item.size(0) == E and key.startswith("edge"):
                data[key] = item[perm]
        data.edge_index = data.edge_index[:, perm]
        value = data[self.attr]
        data.adj_t = _Adj(data.edge_index[1], data.edge_index[0], value, N)
        return data


class Cartesian:
    def __init__(self, norm=True, max_value=None, cat=True):
        self.norm, self.max, self.cat = norm, max_value, cat

    def __call__(self, data):
        (row, col), pos, pseudo = data.edge_index, data.pos, getattr(data, "edge_attr", None)
        cart = pos[row] - pos[col]
        cart = cart.view(-1, 1) if cart.dim() == 1 else cart
        if self.norm and cart.numel() > 0:
            max_value = cart.abs().max() if self.max is None else self.max
            cart = cart / (2 * max_value) + 0.5
        if pseudo is not None and self.cat:
            pseudo = pseudo.view(-1, 1) if pseudo.dim() == 1 else pseudo
            data.edge_attr = torch.cat([pseudo, cart.type_as(pseudo)], dim=-1)
        else:
            data.edge_attr = cart
        return data


class BatchNorm(torch.nn.Module):
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


def _spline_basis(pseudo, kernel_size, is_open_spline, degree):
    from oracle import ref_ops as R
    ks = int(kernel_size[0]) if torch.is_tensor(kernel_size) else int(kernel_size)
    op = bool(is_open_spline[0]) if torch.is_tensor(is_open_spline) else bool(is_open_spline)
    return R.spline_basis(pseudo, ks, op, int(degree))


class SplineConv(torch.nn.Module):
    """torch_geometric.nn.conv.SplineConv (>= 2.0.3 parameter naming: weight, lin.weight, bias)."""

    def __init__(self, in_channels, out_channels, dim, kernel_size, is_open_spline=True, degree=1, aggr="mean",
                 root_weight=True, bias=True, **kwargs):
        super().__init__()
        assert aggr in ("sum", "add"), "the reference configs use aggr=sum"
        self.in_channels, self.out_channels, self.dim, self.degree = in_channels, out_channels, dim, degree
        self.root_weight = root_weight
        self.register_buffer("kernel_size", torch.tensor([kernel_size] * dim, dtype=torch.long))
        self.register_buffer("is_open_spline", torch.tensor([is_open_spline] * dim, dtype=torch.uint8))
        K = kernel_size ** dim
        self.weight = torch.nn.Parameter(torch.randn(K, in_channels, out_channels) * 0.05)
        if root_weight:
            self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        if bias:
            self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def message(self, x_j, edge_attr):
        from oracle import ref_ops as R
        basis, index = _spline_basis(edge_attr, self.kernel_size, self.is_open_spline, self.degree)
        return R.spline_weighting(x_j, self.weight, basis, index)

    def propagate(self, edge_index, x=None, edge_attr=None, size=None):
        adj = edge_index
        x_j = x[0][adj.src]
        msg = self.message(x_j, adj.value)
        out = torch.zeros((adj.n, msg.shape[1]), dtype=msg.dtype)
        out.index_add_(0, adj.dst, msg)                            # segment_csr(sum) in (dst, src) order
        return out


# ------------------------------------------------------------------------------------------------------------------
# torch_scatter / torch_cluster / pool helpers
# ------------------------------------------------------------------------------------------------------------------
def _scatter_max(src, index, dim=0, dim_size=None):
    from oracle import ref_ops as R
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = R.scatter_max(src, index, n)
    return out, None


def _scatter_sum(src, index, dim=0, dim_size=None):
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((n,) + src.shape[1:], dtype=src.dtype)
    return out.index_add_(0, index, src)


def _scatter_mean(src, index, dim=0, dim_size=None):
    from oracle import ref_ops as R
    n = int(index.max()) + 1 if dim_size is None else dim_size
    return R.scatter_mean(src, index, n)


def _avg_pool_x(cluster, x, size=None):
    return _scatter_mean(x, cluster, dim_size=size)


def _pool_pos(cluster, pos):
    return _scatter_mean(pos, cluster)


def _grid_cluster(pos, size, start=None, end=None):
    from oracle import ref_ops as R
    return R.grid_cluster(pos, size, start, end)


# ------------------------------------------------------------------------------------------------------------------
# the radius-graph extension (src/dagr/graph/ev_graph.cu) on the CPU through oracle/graph_oracle.c
# ------------------------------------------------------------------------------------------------------------------
def _insert_in_queue_cuda(sorted_indices, unique_coords, cumsum_counter, queue):
    return queue                                                   # the C oracle keeps its own FIFO


def _fill_edges_cuda(batch, pos, all_timestamps, queue, indices, max_num_neighbors, radius, delta_t_us, edges, min_index):
    from oracle import ref_ops as R
    B, Q, H, W = queue.shape
    g = R.RefGraph(W, H, B, int(max_num_neighbors), Q, int(radius), int(delta_t_us))
    e = g.forward(batch.int(), pos.int())
    edges.fill_(-1)
    edges[:, : e.shape[1]] = e                                     # compaction edges[:, edges[1] >= 0] keeps this order


# ------------------------------------------------------------------------------------------------------------------
# YOLOX (Megvii @618fd8c0): class skeletons the reference subclasses
# ------------------------------------------------------------------------------------------------------------------
class _BaseConv(torch.nn.Module):
    def __init__(self, cin, cout, ksize, stride, act="silu"):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, ksize, stride, (ksize - 1) // 2, bias=False)
        self.bn = torch.nn.BatchNorm2d(cout)
        self.act = torch.nn.SiLU()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class YOLOXHead(torch.nn.Module):
    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024), act="silu", depthwise=False):
        super().__init__()
        self.n_anchors, self.num_classes, self.decode_in_inference = 1, num_classes, True
        for n in ("cls_convs", "reg_convs", "cls_preds", "reg_preds", "obj_preds", "stems"):
            setattr(self, n, torch.nn.ModuleList())
        hid = int(256 * width)
        for c in in_channels:
            self.stems.append(_BaseConv(int(c * width), hid, 1, 1))
            self.cls_convs.append(torch.nn.Sequential(_BaseConv(hid, hid, 3, 1), _BaseConv(hid, hid, 3, 1)))
            self.reg_convs.append(torch.nn.Sequential(_BaseConv(hid, hid, 3, 1), _BaseConv(hid, hid, 3, 1)))
            self.cls_preds.append(torch.nn.Conv2d(hid, num_classes, 1, 1, 0))
            self.reg_preds.append(torch.nn.Conv2d(hid, 4, 1, 1, 0))
            self.obj_preds.append(torch.nn.Conv2d(hid, 1, 1, 1, 0))
        self.strides = strides


class YOLOX(torch.nn.Module):
    def __init__(self, backbone=None, head=None):
        super().__init__()
        self.backbone, self.head = backbone, head

    def forward(self, x, targets=None):
        fpn_outs = self.backbone(x)
        assert not self.training
        return self.head(fpn_outs)


class IOUloss(torch.nn.Module):
    def __init__(self, reduction="none", loss_type="iou"):
        super().__init__()


def install():
    """idempotent; returns the imported reference package root module `dagr`."""
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))                              # oracle/
    for k in [k for k in sys.modules if k == "dagr" or k.startswith("dagr.")]:
        del sys.modules[k]                                         # the repo's own drop-in shim package of the same name
    # the reference's `dagr` is a namespace package (no __init__.py); this repo ships a regular package of the same name
    # (the drop-in shim), which would win the import: bind the name to the reference's directory explicitly
    ref_pkg = types.ModuleType("dagr")
    ref_pkg.__path__ = [REF_SRC + "/dagr"]
    sys.modules["dagr"] = ref_pkg
    _mod("torch_geometric")
    _mod("torch_geometric.data", Data=Data, Batch=Batch)
    _mod("torch_geometric.transforms", Cartesian=Cartesian, ToSparseTensor=ToSparseTensor)
    _mod("torch_geometric.transforms.to_sparse_tensor", ToSparseTensor=ToSparseTensor)
    _mod("torch_geometric.nn", BatchNorm=BatchNorm, SplineConv=SplineConv)
    _mod("torch_geometric.nn.norm", BatchNorm=BatchNorm)
    _mod("torch_geometric.nn.conv", SplineConv=SplineConv)
    _mod("torch_geometric.nn.pool")
    _mod("torch_geometric.nn.pool.avg_pool", _avg_pool_x=_avg_pool_x)
    _mod("torch_geometric.nn.pool.pool", pool_pos=_pool_pos)
    _mod("torch_scatter", scatter_max=_scatter_max, scatter_sum=_scatter_sum, scatter_mean=_scatter_mean)
    _mod("torch_cluster", grid_cluster=_grid_cluster)
    _mod("torch_spline_conv", spline_basis=_spline_basis)
    _mod("yolox")
    _mod("yolox.models", YOLOX=YOLOX, YOLOXHead=YOLOXHead, IOUloss=IOUloss)
    _mod("ev_graph_cuda", insert_in_queue_cuda=_insert_in_queue_cuda, fill_edges_cuda=_fill_edges_cuda,
         insert_in_queue_single_cuda=None)
    _mod("asy_tools", masked_isdiff=None, masked_lin=None, masked_lin_no_bias=None, masked_inplace_BN=None)
    import dagr                                                    # namespace package rooted at /root/reference/src/dagr
    assert any(REF_SRC in str(p) for p in dagr.__path__), dagr.__path__
    return dagr
