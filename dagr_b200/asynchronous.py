"""Incremental (asynchronous) inference: events arrive in chunks, the detector state is updated instead of
recomputed (reference: src/dagr/asynchronous/, driven by evaluate_flops.py:82-165).

What is incremental here, and why it is exact
---------------------------------------------
The event graph is causal: an edge always points from an older to a newer event (ev_graph.cu:64), so
appending events never changes the inputs of an existing node.  Consequently
  * the adjacency and both conv_block1 activations of OLD events are final -- only the NEW events are probed
    and convolved (kernels take `min_idx`; old activations are kept in arrival order and gathered into the new
    cell-major order, asynchronous/conv.py:110-128 "node added" case);
  * pool1 keeps per-voxel running state (channel max, edge-direction mask; positions/counts are re-derived from
    the voxel's events) and only voxels that received events change (max_pool.py:123-154);
  * the coarse levels (<= B*2240 voxels) are recomputed densely from the updated pool1 grid -- on a B200 that is
    cheaper than the reference's change-set bookkeeping (dozens of unique/isin/nonzero host syncs per event).
The reference's correctness criterion is kept and tightened: after any number of steps the outputs equal the
dense forward over all events seen so far (evaluate_flops.py:139-147 uses 1e-3; tests use 1e-5).

Limits (documented): append-only between reset()s -- evicting old events changes the inputs of the
nodes they fed, which needs a re-probe of those nodes; `evict_older_than` therefore rebuilds the live window
with one dense pass.  Events-only model (no image fusion) in streaming mode.
"""
from __future__ import annotations

import torch

from . import _lib
from .data import EventBatch


class StreamState:
    def __init__(self):
        self.cap = 0
        self.xa_arr = None        # f32[cap,16]  conv_block1.conv_block1 activations in arrival order
        self.voxmax = None        # f32[cells,16] running per-voxel channel max (pool1)
        self.cellmask = None      # i32[cells]   running coarse in-edge mask of pool1 voxels
        self.n = 0
        self._geom_id = None

    def ensure(self, geom, N, dev):
        if self._geom_id != (geom.W, geom.H, geom.B, geom.r):
            self.voxmax = torch.full((geom.cells1, 16), float("-inf"), dtype=torch.float32, device=dev)
            self.cellmask = torch.zeros(geom.cells1, dtype=torch.int32, device=dev)
            self._geom_id = (geom.W, geom.H, geom.B, geom.r)
            self.cap = 0
        if N > self.cap:
            cap = max(int(N * 1.5), 4096)
            new = torch.empty((cap, 16), dtype=torch.float32, device=dev)
            if self.xa_arr is not None and self.n > 0:
                new[: self.n] = self.xa_arr[: self.n]
            self.xa_arr, self.cap = new, cap

    def reset(self):
        self.n = 0
        if self.voxmax is not None:
            self.voxmax.fill_(float("-inf"))
            self.cellmask.zero_()


class AsyncDAGR:
    """stateful wrapper: `step(chunk)` appends events and returns the detections for everything seen so far."""

    def __init__(self, model):
        if model.backbone.use_image:
            raise NotImplementedError("streaming mode supports the events-only model")
        self.model = model
        self.state = StreamState()
        self._batch = self._pos = self._feat = None
        self._hb = self._hp = self._hf = None
        self._n = 0
        self.B = self.W = self.H = None

    def reset(self):
        self.state.reset()
        self._batch = self._pos = self._feat = None
        self._hb = self._hp = self._hf = None
        self._n = 0

    @property
    def num_events(self):
        return 0 if self._batch is None else int(self._batch.shape[0])

    @torch.no_grad()
    def step_decoded(self, chunk: EventBatch, batch_size=None):
        """chunk: formatted EventBatch (same contract as DAGR.forward); events of a sample must be newer than the
        ones already seen for that sample.  Returns decoded head outputs [B, A, 5+nc]."""
        m = self.model
        batch_i, pos_i, feat, W, H = m._prepare_events(chunk)
        B = int(batch_size or getattr(chunk, "num_graphs", 1) or 1)
        k = int(batch_i.shape[0])
        if self._hb is None:
            self.B, self.W, self.H = B, W, H
            self._n = 0
        else:
            assert (B, W, H) == (self.B, self.W, self.H), "stream geometry changed; call reset()"
        if self._hb is None or self._n + k > self._hb.shape[0]:          # history buffers grow geometrically
            cap = max(2 * (self._n + k), 65536)
            dev = batch_i.device
            hb, hp, hf = (torch.empty(cap, dtype=torch.int32, device=dev), torch.empty((cap, 3), dtype=torch.int32, device=dev),
                          torch.empty(cap, dtype=torch.float32, device=dev))
            if self._hb is not None and self._n:
                hb[: self._n] = self._hb[: self._n]; hp[: self._n] = self._hp[: self._n]; hf[: self._n] = self._hf[: self._n]
            self._hb, self._hp, self._hf = hb, hp, hf
        n0 = self._n
        self._hb[n0:n0 + k] = batch_i; self._hp[n0:n0 + k] = pos_i; self._hf[n0:n0 + k] = feat
        self._n = n0 + k
        self._batch, self._pos, self._feat = self._hb[: self._n], self._hp[: self._n], self._hf[: self._n]
        n_old = self.state.n
        dec = m.engine.forward_events(self._batch, self._pos, self._feat, self.B, self.W, self.H, stream_state=self.state, n_old=n_old)
        self.state.n = self.num_events
        return dec

    @torch.no_grad()
    def step(self, chunk: EventBatch, batch_size=None, filtering=True):
        m = self.model
        dec = self.step_decoded(chunk, batch_size)
        det, ndet = m.engine.postprocess(dec, m.conf_threshold, m.nms_threshold, m.width, m.height, filtering=filtering)
        out = []
        for b, n in enumerate(ndet.tolist()):
            d = det[b, :n]
            out.append(dict(boxes=d[:, :4], scores=d[:, 4], labels=d[:, 5].long()))
        return out

    @torch.no_grad()
    def evict_older_than(self, t_us: int):
        """sliding window: drop events with t < t_us and rebuild the state with one dense pass over the live window."""
        if self._batch is None:
            return
        keep = self._pos[:, 2] >= int(t_us)
        self._batch, self._pos, self._feat = self._batch[keep].contiguous(), self._pos[keep].contiguous(), self._feat[keep].contiguous()
        self._hb, self._hp, self._hf, self._n = self._batch, self._pos, self._feat, int(self._batch.shape[0])
        self.state.reset()
        dec = self.model.engine.forward_events(self._batch, self._pos, self._feat, self.B, self.W, self.H, stream_state=self.state,
                                               n_old=0)
        self.state.n = self.num_events
        return dec


def make_model_asynchronous(model, log_flops: bool = False):
    """name-compatible entry point (src/dagr/asynchronous/__init__.py:41): returns the stateful wrapper."""
    return AsyncDAGR(model)
