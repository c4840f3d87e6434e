"""Minimal re-statement of the YOLOX (Megvii, @618fd8c0) module layout the reference inherits from
(src/dagr/model/networks/dagr.py:6,14,106,125): only what is needed for state_dict-key compatibility
and for the dense CNN head (cuDNN path, tensor cores allowed there).  YOLOX itself is not installed
offline; losses (training) are out of scope.
"""
from __future__ import annotations

import torch
from torch import nn


class BaseConv(nn.Module):
    """Conv2d(bias=False) -> BatchNorm2d -> SiLU."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="silu"):
        super().__init__()
        pad = (ksize - 1) // 2
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=ksize, stride=stride, padding=pad,
                              groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.SiLU(inplace=True) if act == "silu" else nn.ReLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class YOLOXHead(nn.Module):
    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024), act="silu",
                 depthwise=False):
        super().__init__()
        self.n_anchors = 1
        self.num_classes = num_classes
        self.decode_in_inference = True
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        self.cls_preds = nn.ModuleList()
        self.reg_preds = nn.ModuleList()
        self.obj_preds = nn.ModuleList()
        self.stems = nn.ModuleList()
        hid = int(256 * width)
        for i in range(len(in_channels)):
            self.stems.append(BaseConv(int(in_channels[i] * width), hid, ksize=1, stride=1, act=act))
            self.cls_convs.append(nn.Sequential(BaseConv(hid, hid, 3, 1, act=act), BaseConv(hid, hid, 3, 1, act=act)))
            self.reg_convs.append(nn.Sequential(BaseConv(hid, hid, 3, 1, act=act), BaseConv(hid, hid, 3, 1, act=act)))
            self.cls_preds.append(nn.Conv2d(hid, self.n_anchors * self.num_classes, 1, 1, 0))
            self.reg_preds.append(nn.Conv2d(hid, 4, 1, 1, 0))
            self.obj_preds.append(nn.Conv2d(hid, self.n_anchors * 1, 1, 1, 0))
        self.strides = strides


class YOLOX(nn.Module):
    def __init__(self, backbone=None, head=None):
        super().__init__()
        self.backbone = backbone
        self.head = head
