"""oracle/ref_model.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's synchronous forward (events-only and image-fusion variants):
Net.forward (src/dagr/model/networks/net.py:108-190), GNNHead.forward/process_feature/
collect_outputs/decode_outputs (dagr.py:179-312), postprocess (model/utils.py:61-110), driven by a
state_dict with the reference's key names.  Built from oracle/ref_ops.py; parity status: see the
header of ref_ops.py (graph build pinned by the reference's own CUDA kernels; PyG/YOLOX-derived
ops PARITY UNPINNED -- no golden vectors exist upstream).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ref_ops as R


def _bn(sd, prefix, x):
    return R.batch_norm_eval(x, sd[prefix + ".module.weight"], sd[prefix + ".module.bias"],
                             sd[prefix + ".module.running_mean"], sd[prefix + ".module.running_var"], 1e-5)


class _Graph:
    def __init__(self, x, pos, batch, edge_index, edge_attr):
        self.x, self.pos, self.batch, self.edge_index, self.edge_attr = x, pos, batch, edge_index, edge_attr

    def copy(self):
        return _Graph(self.x.clone(), self.pos, self.batch, self.edge_index, self.edge_attr)


class RefModel:
    def __init__(self, sd: Dict[str, torch.Tensor], args, height: int, width: int, use_lut: bool = True):
        self.sd = {k: v.detach().cpu().float() if v.dtype.is_floating_point else v.detach().cpu() for k, v in sd.items()}
        self.args, self.H, self.W = args, int(height), int(width)
        self.use_lut = use_lut
        self.poolings = R.compute_pooling_at_each_layer(args.pooling_dim_at_output, 4)
        self.max_vals = 2 * self.poolings[:, :2].max(-1).values
        self.effective_radius = 2 * float(int(args.radius * width + 2) / width)
        self.strides = torch.ceil(self.poolings[-2:, 1] * height).numpy().astype("int32").tolist()[-args.num_scales:]
        self.num_classes = dict(dsec=2, ncaltech101=100).get(args.dataset, 2)
        self.graph = None
        self._luts = {}

    # --- one SplineConv with the LUT the reference would have cached for it (dagr.py:37-72) ------
    def _conv(self, prefix, g: _Graph, lut_level):
        sd = self.sd
        w = sd[prefix + ".weight"]
        root = sd[prefix + ".lin.weight"]
        bias = sd.get(prefix + ".bias")
        lut = remap = None
        if self.use_lut and g.edge_index.shape[1] > 0:
            # evaluate the LUT only at the offsets that occur (the full LUT is GBs at coarse levels)
            rx, ry, M = lut_level
            lut, remap = self._lazy_lut(prefix, w, rx, ry, M)
        return R.spline_conv(g.x, g.edge_index, g.edge_attr, w, root, bias, lut=lut, remap=remap)

    def _lazy_lut(self, prefix, w, rx, ry, M):
        key = (prefix, rx, ry)
        if key not in self._luts:
            cin, cout = w.shape[1:]
            if (2 * rx + 1) * (2 * ry + 1) * cin * cout > 60_000_000:
                self._luts[key] = _SparseLut(w, self.H, self.W, rx, M, ry, M)
            else:
                lut, remap = R.build_lut(w, self.H, self.W, rx, M, ry, M)
                self._luts[key] = (lut, remap)
        v = self._luts[key]
        return v if isinstance(v, tuple) else (v, v.remap)

    def _layer(self, prefix, g: _Graph, lut_level):
        """Layer.forward (conv.py:59-72)."""
        sd = self.sd
        skip_x = g.x.clone()
        g = g.copy()
        g.x = torch.relu(_bn(sd, prefix + ".conv_block1.norm", self._conv(prefix + ".conv_block1.conv", g, lut_level)))
        xa = g.x.clone()
        y = _bn(sd, prefix + ".conv_block2.norm", self._conv(prefix + ".conv_block2.conv", g, lut_level))
        sk = _bn(sd, prefix + ".conv_block2.norm_skip", skip_x @ sd[prefix + ".conv_block2.lin.mlp.weight"].t())
        g.x = torch.relu(y + sk)
        return g, xa

    def _lut_level(self, i):
        """(rx, ry, M) as DAGR.cache_luts derives them: i = 0 event level, 1..4 after pool i."""
        if i == 0:
            r = int(self.args.radius * self.W + 1)
            return r, r, self.effective_radius
        vs = self.poolings[i - 1]
        rx = int(np.ceil(2 * vs[0].numpy() * self.W))
        ry = int(np.ceil(2 * vs[1].numpy() * self.H))
        M = 2 * self.effective_radius if i == 1 else self.max_vals[i - 1]
        return rx, ry, M

    # --------------------------------------------------------------------------------------------
    def forward(self, x, pos, batch, batch_size, image_feats=None, image_outs=None, filtering=True,
                conf_thre=0.001, nms_thre=0.65, pos_hints=None, mirror_t_quirk=True):
        """x fp32[N,1], pos fp32[N,3] normalised, batch int64[N]  (all CPU).
        pos_hints: optional list of 4 fp32 [n_level,2] tensors, see ref_ops.pooling(pos_hint=...)."""
        args, W, H = self.args, self.W, self.H
        out = {}
        T = int(getattr(args, "time_window_us", 1000000))
        # EV_TGN (ev_tgn.py:39-59)
        pos_i = R.denormalize_pos(pos, W, H, T)
        if self.graph is None or self.graph.B != batch_size:
            self.graph = R.RefGraph(W, H, batch_size, args.max_neighbors, 128, int(args.radius * W + 1),
                                    int(args.radius * T))
        else:
            self.graph.reset()
        edge_index = self.graph.forward(batch.int(), pos_i, delete_nodes=False, collect_edges=True)
        out["edge_index"] = edge_index
        g = _Graph(x.float(), pos, batch, edge_index, None)
        if image_feats is not None:
            g.x = torch.cat((g.x, R.sample_features(g.pos, g.batch, image_feats[0], W, H)), dim=1)
        # Cartesian + clamp (net.py:122-123)
        g.edge_attr = torch.clamp(R.cartesian(pos, edge_index, self.effective_radius), min=0, max=1)
        g.x = torch.cat((g.x, pos[:, :2]), dim=1)
        g, xa = self._layer("backbone.conv_block1", g, self._lut_level(0))
        out["x1a"], out["x1"] = xa, g.x.clone()
        levels = []
        cart_max = [2 * self.effective_radius, self.max_vals[1], self.max_vals[2], self.max_vals[3]]
        aggrs = [args.pooling_aggr, args.pooling_aggr, args.pooling_aggr, "mean"]
        outs = []
        for i in range(4):
            if image_feats is not None:
                g.x = torch.cat((g.x, R.sample_features(g.pos, g.batch, image_feats[i + 1], W, H)), dim=1)
            if g.x.shape[0] == 0:
                pooled = None
            else:
                pooled = R.pooling(g.x, g.pos, g.batch, g.edge_index, self.poolings[i], W, H, batch_size, cart_max[i],
                                   aggr=aggrs[i], keep_temporal_ordering=getattr(args, "keep_temporal_ordering", False),
                                   pos_hint=None if pos_hints is None else pos_hints[i], mirror_t_quirk=mirror_t_quirk)
            if pooled is None:
                levels.append(None)
                pg = g
            else:
                levels.append(pooled)
                pg = _Graph(pooled["x"], pooled["pos"], pooled["batch"], pooled["edge_index"], pooled["edge_attr"])
            pg.x = torch.cat((pg.x, pg.pos[:, :2]), dim=1)
            g, _ = self._layer(f"backbone.layer{i + 2}", pg, self._lut_level(i + 1))
            if i >= 2:
                outs.append((g.copy(), self.poolings[i][:3], i + 1))
        out["levels"] = levels
        out["out3"], out["out4"] = outs[0][0].x, outs[1][0].x
        # head (dagr.py:192-236)
        scales = outs[-args.num_scales:]
        dense_outs, hw = [], []
        for k, (gk, pooling3, lut_i) in enumerate(scales):
            sfx = str(k + 1)
            lv = self._lut_level(lut_i)
            sd = self.sd

            def block(prefix, gin):
                gg = gin.copy()
                gg.x = torch.relu(_bn(sd, prefix + ".norm", self._conv(prefix + ".conv", gg, lv)))
                return gg

            stem = block("head.stem" + sfx, gk)
            cls_feat = block("head.cls_conv" + sfx, stem)
            reg_feat = block("head.reg_conv" + sfx, stem)
            d = {}
            for name, feat in (("cls", cls_feat), ("reg", reg_feat), ("obj", reg_feat)):
                y = self._conv(f"head.{name}_pred{sfx}", feat, lv)
                d[name] = R.to_dense(y, feat.pos, pooling3, feat.batch, batch_size)
                if image_outs is not None:
                    d[name] = d[name] + image_outs[name + "_output"][k]
            dense_outs.append(d)
            o = torch.cat([d["reg"], torch.sigmoid(d["obj"]), torch.sigmoid(d["cls"])], 1)   # dagr.py:300-302
            hw.append(o.shape[-2:])
            d["cat"] = o
        out["dense"] = dense_outs
        outputs = torch.cat([d["cat"].flatten(start_dim=2) for d in dense_outs], dim=2).permute(0, 2, 1)
        decoded = R.decode_outputs(outputs, hw, self.strides)
        out["decoded"] = decoded
        out["detections"] = R.postprocess_network_output(decoded, self.num_classes, conf_thre, nms_thre, height=H, width=W,
                                                         filtering=filtering)
        return out


class _SparseLut:
    """message_lut for LUTs too large to materialise: evaluates init_lut's entry on demand
    (identical arithmetic to build_lut for the entries that are touched)."""

    def __init__(self, weight, height, width, rx, Mx, ry, My):
        self.weight, self.h, self.w, self.rx, self.ry, self.Mx, self.My = weight, height, width, rx, ry, Mx, My
        self.remap = torch.Tensor([[2 * Mx * width, 0, -Mx * width + rx], [0, 2 * My * height, -My * height + ry]])

    def __getitem__(self, idx):
        dxi, dyi = idx
        dx = (dxi - self.rx).float() / (2 * self.Mx * self.w) + 0.5
        dy = (dyi - self.ry).float() / (2 * self.My * self.h) + 0.5
        attr = torch.stack([dx, dy], dim=1)
        bil_w, indices = R.spline_basis(attr, 5, True, 1)
        return (bil_w[..., None, None] * self.weight[indices]).sum(1)
