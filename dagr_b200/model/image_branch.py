"""The dense image branch of the hybrid detector -- ResNet trunk taps (`HookModule`, reference
src/dagr/model/networks/net_img.py:42-135), nearest resize of the two output taps and the YOLOX-style `CNNHead`
(dagr.py:106-122,205-206) -- as ONE replayed CUDA graph on a side stream.

The branch does not depend on the events, so it runs concurrently with the event-level graph kernels (cell-major sort +
radius-graph probe) of the same forward; the engine waits for its event right before the first kernel that samples a
feature map.  Arithmetic is untouched: the same torch modules run (cuDNN, TF32 convolutions like the reference's default),
only captured once per input shape instead of ~250 eager launches per forward -- at batch 1 the eager trunk is bound by
host launch time, not by the GPU.
"""
from __future__ import annotations

import torch


class ImageBranch:
    def __init__(self, model):
        self.model = model
        self.stream = None
        self._sizes = None         # head grid sizes: read once (a device->host read is not allowed under graph capture)
        self._graphs = {}          # (B, C, H, W, device) -> dict(graph, inp, feats, outs, warm)

    def invalidate(self):
        self._graphs = {}

    def _compute(self, image):
        m = self.model
        feats, outs = m.backbone.net(image)
        feats = [f.float().contiguous() for f in feats]
        if self._sizes is None:
            self._sizes = m.backbone.get_output_sizes()[-m.head.num_scales:]
        sizes = self._sizes
        cnn_in = [torch.nn.functional.interpolate(o, size=tuple(sz)) for o, sz in zip(outs[-m.head.num_scales:], sizes)]
        image_outs = m.head.cnn_head(cnn_in)
        return feats, {k: [t.float().contiguous() for t in v] for k, v in image_outs.items()}

    @torch.no_grad()
    def run(self, image: torch.Tensor, use_graph: bool = True):
        """-> (image_feats, image_outs, event): tensors are valid on any stream that waited for `event`; they are
        overwritten by the next call (static buffers of the captured graph)."""
        dev = image.device
        cur = torch.cuda.current_stream(dev)
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        s = self.stream
        s.wait_stream(cur)                                   # the image is ready and every consumer of the previous outputs is done
        key = (tuple(image.shape), str(dev))
        st = self._graphs.get(key)
        with torch.cuda.stream(s):
            if not use_graph:
                feats, outs = self._compute(image.float())
            elif st is None or st["graph"] is None:
                if st is None:
                    st = dict(graph=None, inp=torch.empty(image.shape, dtype=torch.float32, device=dev), warm=0)
                    self._graphs[key] = st
                st["inp"].copy_(image)
                feats, outs = self._compute(st["inp"])      # eager warm-up (cuDNN algorithm selection, workspaces)
                st["warm"] += 1
                if st["warm"] >= 2:
                    s.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        gf, go = self._compute(st["inp"])
                    st.update(graph=g, feats=gf, outs=go)
                    g.replay()                               # fill the static outputs for this call
                    feats, outs = gf, go
            else:
                st["inp"].copy_(image)
                st["graph"].replay()
                feats, outs = st["feats"], st["outs"]
            ev = torch.cuda.Event()
            ev.record(s)
        for t in feats:
            t.record_stream(cur)
        for v in outs.values():
            for t in v:
                t.record_stream(cur)
        return feats, outs, ev
