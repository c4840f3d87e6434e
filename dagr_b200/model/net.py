"""Backbone container (reference: src/dagr/model/networks/net.py:31-106): same sub-module names and
shapes.  The forward lives in dagr_b200.engine."""
from __future__ import annotations

import torch
from torch import nn

from ..geometry import compute_pooling_at_each_layer
from .layers import Cartesian, EV_TGN, Layer, Pooling
from .net_img import HookModule


def _make_img_net(name: str):
    import torchvision
    fn = getattr(torchvision.models, name)
    try:
        return fn(weights=None)         # no network: random init (reference uses pretrained=True, net.py:43)
    except TypeError:
        return fn(pretrained=False)


class Net(nn.Module):
    def __init__(self, args, height, width):
        super().__init__()
        channels = [1, int(args.base_width * 32), int(args.after_pool_width * 64),
                    int(args.net_stem_width * 128), int(args.net_stem_width * 128), int(args.net_stem_width * 128)]
        self.out_channels_cnn = []
        self.use_image = bool(getattr(args, "use_image", False))
        if self.use_image:
            self.out_channels_cnn = [256, 256]
            self.net = HookModule(_make_img_net(args.img_net), input_channels=3, height=height, width=width,
                                  feature_layers=["conv1", "layer1", "layer2", "layer3", "layer4"],
                                  output_layers=["layer3", "layer4"], feature_channels=channels[1:],
                                  output_channels=self.out_channels_cnn)
        self.num_scales = args.num_scales
        self.num_classes = dict(dsec=2, ncaltech101=100).get(args.dataset, 2)
        self.events_to_graph = EV_TGN(args)
        output_channels = channels[1:]
        self.out_channels = output_channels[-2:]
        input_channels = channels[:-1]
        if self.use_image:
            input_channels = [input_channels[i] + self.net.feature_channels[i] for i in range(len(input_channels))]
        self.input_channels = input_channels
        self.output_channels = output_channels

        poolings = compute_pooling_at_each_layer(args.pooling_dim_at_output, num_layers=4)
        self.poolings = poolings
        max_vals = 2 * poolings[:, :2].max(-1).values
        self.strides = torch.ceil(poolings[-2:, 1] * height).numpy().astype("int32").tolist()
        self.strides = self.strides[-self.num_scales:]
        effective_radius = 2 * float(int(args.radius * width + 2) / width)
        self.edge_attrs = Cartesian(norm=True, cat=False, max_value=effective_radius)
        kto = bool(getattr(args, "keep_temporal_ordering", False))
        self.conv_block1 = Layer(2 + input_channels[0], output_channels[0], args=args)
        self.pool1 = Pooling(poolings[0], width=width, height=height, batch_size=args.batch_size,
                             transform=Cartesian(True, 2 * effective_radius), aggr=args.pooling_aggr,
                             keep_temporal_ordering=kto)
        self.layer2 = Layer(input_channels[1] + 2, output_channels[1], args=args)
        self.pool2 = Pooling(poolings[1], width=width, height=height, batch_size=args.batch_size,
                             transform=Cartesian(True, max_vals[1]), aggr=args.pooling_aggr, keep_temporal_ordering=kto)
        self.layer3 = Layer(input_channels[2] + 2, output_channels[2], args=args)
        self.pool3 = Pooling(poolings[2], width=width, height=height, batch_size=args.batch_size,
                             transform=Cartesian(True, max_vals[2]), aggr=args.pooling_aggr, keep_temporal_ordering=kto)
        self.layer4 = Layer(input_channels[3] + 2, output_channels[3], args=args)
        self.pool4 = Pooling(poolings[3], width=width, height=height, batch_size=args.batch_size,
                             transform=Cartesian(True, max_vals[3]), aggr="mean", keep_temporal_ordering=kto)   # net.py:96-97
        self.layer5 = Layer(input_channels[4] + 2, output_channels[4], args=args)

    def get_output_sizes(self):
        poolings = [self.pool3.voxel_size[:2], self.pool4.voxel_size[:2]]
        return [(1 / p + 1e-3).cpu().int().numpy().tolist()[::-1] for p in poolings]
