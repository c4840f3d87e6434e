"""Parameter containers mirroring the reference's graph layers (same attribute / state_dict names).

reference: src/dagr/model/layers/{spline_conv,conv,components,pooling,ev_tgn}.py.  The arithmetic
is not here: `dagr_b200.engine.Engine` reads these modules' tensors and drives the CUDA kernels.
Key layout follows PyG >= 2.0.3 naming: SplineConv.{weight, lin.weight, bias},
BatchNorm.module.*, Linear.mlp.*.
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import nn


class _Lin(nn.Module):
    """torch_geometric.nn.dense.linear.Linear(bias=False): holds `weight` [out, in]."""

    def __init__(self, ic, oc):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(oc, ic))
        bound = 1.0 / math.sqrt(max(ic, 1))
        nn.init.uniform_(self.weight, -bound, bound)


class MySplineConv(nn.Module):
    """SplineConv(dim=2, kernel_size=5, degree=1, open splines, aggr=sum, root weight)
    (spline_conv.py:9-14)."""

    def __init__(self, in_channels, out_channels, args, bias=False, degree=1):
        super().__init__()
        assert degree == 1
        self.in_channels, self.out_channels = in_channels, out_channels
        self.dim = getattr(args, "edge_attr_dim", 2)
        ks = getattr(args, "kernel_size", 5)
        if self.dim != 2 or getattr(args, "aggr", "sum") not in ("sum", "add"):
            raise ValueError("dagr_b200 supports edge_attr_dim=2, aggr=sum (the reference configs)")
        K = ks ** self.dim
        self.weight = nn.Parameter(torch.empty(K, in_channels, out_channels))
        self.lin = _Lin(in_channels, out_channels)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.register_buffer("kernel_size", torch.tensor([ks] * self.dim, dtype=torch.long))
        self.register_buffer("is_open_spline", torch.tensor([1] * self.dim, dtype=torch.uint8))
        bound = 1.0 / math.sqrt(in_channels * K)
        nn.init.uniform_(self.weight, -bound, bound)
        self.lut_params = None

    def init_lut(self, height, width, rx=None, Mx=None, ry=None, My=None):
        """spline_conv.py:16-37.  The reference materialises lut[2rx+1, 2ry+1, Cin, Cout] (GBs at the
        coarse levels); our kernels evaluate the identical basis at the integer offsets on
        the fly, so only the parameters are recorded."""
        self.lut_params = dict(height=height, width=width, rx=rx, Mx=Mx, ry=ry or rx, My=My or Mx)


class SplineConvToDense(MySplineConv):
    pass


class BatchNormData(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps=1e-5, momentum=0.1)


class Linear(nn.Module):
    def __init__(self, ic, oc, bias=True):
        super().__init__()
        self.mlp = nn.Linear(ic, oc, bias=bias)


class Cartesian(nn.Module):
    def __init__(self, norm=True, max_value=None, cat=False):
        super().__init__()
        self.norm, self.max, self.cat = norm, max_value, cat


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, args, degree=1):
        super().__init__()
        self.activation_name = getattr(args, "activation", "relu")
        self.conv = MySplineConv(in_channels, out_channels, args=args, bias=False, degree=degree)
        self.norm = BatchNormData(out_channels)


class ConvBlockWithSkip(nn.Module):
    def __init__(self, in_channel, out_channel, skip_in_channel, args):
        super().__init__()
        self.activation_name = getattr(args, "activation", "relu")
        self.conv = MySplineConv(in_channel, out_channel, args=args, bias=False)
        self.norm = BatchNormData(out_channel)
        self.lin = Linear(skip_in_channel, out_channel, bias=False)
        self.norm_skip = BatchNormData(out_channel)


class Layer(nn.Module):
    def __init__(self, in_channels, out_channels, args):
        super().__init__()
        self.in_channel, self.out_channel = in_channels, out_channels
        self.conv_block1 = ConvBlock(in_channels, out_channels, args)
        self.conv_block2 = ConvBlockWithSkip(out_channels, out_channels, in_channels, args=args)


class Pooling(nn.Module):
    """pooling.py:19-45: only non-persistent buffers, no parameters."""

    def __init__(self, size, width, height, batch_size, transform, aggr="max", keep_temporal_ordering=False):
        super().__init__()
        assert aggr in ("mean", "max")
        self.aggr = aggr
        self.register_buffer("voxel_size", torch.cat([size, torch.Tensor([1])]), persistent=False)
        self.transform = transform
        self.keep_temporal_ordering = keep_temporal_ordering
        self.register_buffer("wh_inv", 1 / torch.Tensor([[width, height]]), persistent=False)

    @property
    def num_grid_cells(self):
        return (1 / self.voxel_size + 1e-3).int().prod()


class EV_TGN(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.radius = args.radius
        self.max_neighbors = args.max_neighbors
        self.max_queue_size = 128          # ev_tgn.py:24
