"""The dense image branch of the hybrid detector -- ResNet trunk taps (`HookModule`, reference
src/dagr/model/networks/net_img.py:42-135), nearest resize of the two output taps and the YOLOX-style `CNNHead`
(dagr.py:106-122,205-206) -- as TWO replayed CUDA graphs on a side stream.

The branch does not depend on the events, so it runs concurrently with the graph kernels of the same forward:
  stage 1 (conv1 .. layer1 + their 1x1 tap convs) next to the cell-major sort and the radius-graph probe; the event-level
          convs wait for its event (they sample the conv1 / layer1 taps);
  stage 2 (layer2 .. layer4, remaining tap convs, resize, CNN head) next to the event-level convs (latency-bound fp32 SIMT
          kernels at ~30 % occupancy, while the trunk is tensor-core work); the coarse stack waits for its event.
Arithmetic is untouched: the same torch modules run (cuDNN, TF32 convolutions like the reference's default), only captured
once per input shape instead of ~250 eager launches per forward -- at batch 1 the eager trunk is bound by host launch
time, not by the GPU.
"""
from __future__ import annotations

import torch


class ImageBranch:
    def __init__(self, model):
        self.model = model
        self.stream = None
        self._sizes = None         # head grid sizes: read once (a device->host read is not allowed under graph capture)
        self._graphs = {}          # (shape, device) -> dict(g1, g2, inp, mid, feats, outs, warm)

    def invalidate(self):
        self._graphs = {}

    def _stage1(self, image):
        taps, mid = self.model.backbone.net.stage1(image)
        return [t.float().contiguous() for t in taps], mid

    def _stage2(self, mid):
        m = self.model
        feats, outs = m.backbone.net.stage2(mid)
        if self._sizes is None:
            self._sizes = m.backbone.get_output_sizes()[-m.head.num_scales:]
        cnn_in = [torch.nn.functional.interpolate(o, size=tuple(sz)) for o, sz in zip(outs[-m.head.num_scales:], self._sizes)]
        image_outs = m.head.cnn_head(cnn_in)
        return [f.float().contiguous() for f in feats], {k: [t.float().contiguous() for t in v] for k, v in image_outs.items()}

    @torch.no_grad()
    def run(self, image: torch.Tensor, use_graph: bool = True):
        """-> (image_feats, image_outs, (event1, event2)): the first two feature maps are valid on any stream that waited
        for event1, everything else after event2; all of them are overwritten by the next call (static graph buffers)."""
        dev = image.device
        cur = torch.cuda.current_stream(dev)
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        s = self.stream
        s.wait_stream(cur)                                   # the image is ready and every consumer of the previous outputs is done
        key = (tuple(image.shape), str(dev))
        st = self._graphs.get(key)
        with torch.cuda.stream(s):
            ev1, ev2 = torch.cuda.Event(), torch.cuda.Event()
            if not use_graph:
                f12, mid = self._stage1(image.float())
                ev1.record(s)
                f345, outs = self._stage2(mid)
            elif st is None or st.get("g1") is None:
                if st is None:
                    st = dict(g1=None, inp=torch.empty(image.shape, dtype=torch.float32, device=dev), warm=0)
                    self._graphs[key] = st
                st["inp"].copy_(image)
                f12, mid = self._stage1(st["inp"])           # eager warm-up (cuDNN algorithm selection, workspaces)
                ev1.record(s)
                f345, outs = self._stage2(mid)
                st["warm"] += 1
                if st["warm"] >= 2:
                    s.synchronize()
                    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1, stream=s):
                        gf12, gmid = self._stage1(st["inp"])
                    with torch.cuda.graph(g2, stream=s, pool=g1.pool()):
                        gf345, gouts = self._stage2(gmid)
                    st.update(g1=g1, g2=g2, f12=gf12, f345=gf345, outs=gouts)
                    g1.replay()                              # fill the static outputs for this call
                    ev1 = torch.cuda.Event()
                    ev1.record(s)
                    g2.replay()
                    f12, f345, outs = gf12, gf345, gouts
            else:
                st["inp"].copy_(image)
                st["g1"].replay()
                ev1.record(s)
                st["g2"].replay()
                f12, f345, outs = st["f12"], st["f345"], st["outs"]
            ev2.record(s)
        feats = list(f12) + list(f345)
        for t in feats:
            t.record_stream(cur)
        for v in outs.values():
            for t in v:
                t.record_stream(cur)
        return feats, outs, (ev1, ev2)
