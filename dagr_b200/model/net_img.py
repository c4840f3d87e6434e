"""Image trunk (reference: src/dagr/model/networks/net_img.py:42-135).  Dense convs stay torch/cuDNN
(north_star: tensor cores only here).  Same parameter names (`module.*`, `feature_dconv.*`,
`output_dconv.*`); the forward taps conv1 (pre-BN), layer1..4 explicitly instead of through forward
hooks and skips the unused avgpool+fc."""
from __future__ import annotations

import torch
from torch import nn


class HookModule(nn.Module):
    def __init__(self, module, height, width, input_channels=3, feature_layers=(), output_layers=(),
                 feature_channels=None, output_channels=None):
        super().__init__()
        self.module = module
        self.feature_layers = list(feature_layers)
        self.output_layers = list(output_layers)
        with torch.no_grad():
            feats, outs = self._trunk(torch.zeros(1, input_channels, height, width))
        self.feature_channels = [f.shape[1] for f in feats]
        self.output_channels = [o.shape[1] for o in outs]
        self.feature_dconv = nn.ModuleList()
        if feature_channels is not None:
            assert len(feature_channels) == len(self.feature_channels)
            self.feature_dconv = nn.ModuleList(
                [nn.Conv2d(cin, cout, kernel_size=1) for cin, cout in zip(self.feature_channels, feature_channels)])
            self.feature_channels = list(feature_channels)
        self.output_dconv = nn.ModuleList()
        if output_channels is not None:
            assert len(output_channels) == len(self.output_channels)
            self.output_dconv = nn.ModuleList(
                [nn.Conv2d(cin, cout, kernel_size=1) for cin, cout in zip(self.output_channels, output_channels)])
            self.output_channels = list(output_channels)

    # kept for ModelEMA (src/dagr/model/networks/ema.py:25-33)
    def remove_hooks(self):
        pass

    def register_hooks(self):
        pass

    def _trunk(self, x):
        m = self.module
        taps = {}
        x = m.conv1(x); taps["conv1"] = x
        x = m.maxpool(m.relu(m.bn1(x)))
        x = m.layer1(x); taps["layer1"] = x
        x = m.layer2(x); taps["layer2"] = x
        x = m.layer3(x); taps["layer3"] = x
        x = m.layer4(x); taps["layer4"] = x
        return [taps[l] for l in self.feature_layers], [taps[l] for l in self.output_layers]

    def forward(self, x):
        feats, outs = self._trunk(x)
        if len(self.feature_dconv) > 0:
            feats = [d(f) for f, d in zip(feats, self.feature_dconv)]
        if len(self.output_dconv) > 0:
            outs = [d(o) for o, d in zip(outs, self.output_dconv)]
        return feats, outs

    # The same forward in two stages (dagr_b200/model/image_branch.py): the event-level kernels only sample the conv1 and
    # layer1 taps, so they can start while layer2..4 are still running.
    def stage1(self, x):
        """-> (feature maps of the taps up to layer1, trunk activation after layer1)"""
        assert self.feature_layers[:2] == ["conv1", "layer1"] and not set(self.output_layers) & {"conv1", "layer1"}
        m = self.module
        c1 = m.conv1(x)
        l1 = m.layer1(m.maxpool(m.relu(m.bn1(c1))))
        taps = [c1, l1]
        if len(self.feature_dconv) > 0:
            taps = [self.feature_dconv[i](t) for i, t in enumerate(taps)]
        return taps, l1

    def stage2(self, l1):
        """-> (feature maps of the remaining taps, output taps)"""
        m = self.module
        taps = {}
        x = m.layer2(l1); taps["layer2"] = x
        x = m.layer3(x); taps["layer3"] = x
        x = m.layer4(x); taps["layer4"] = x
        feats = [taps[l] for l in self.feature_layers[2:]]
        outs = [taps[l] for l in self.output_layers]
        if len(self.feature_dconv) > 0:
            feats = [self.feature_dconv[2 + i](f) for i, f in enumerate(feats)]
        if len(self.output_dconv) > 0:
            outs = [d(o) for o, d in zip(outs, self.output_dconv)]
        return feats, outs
