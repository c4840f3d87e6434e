"""Multi-GPU: the batch dimension shards cleanly (per-sample FIFOs, batch is a clustering dimension,
eval-mode BN).  One process per GPU; the only collective is one all_gather of the
fixed-size, padded post-NMS detections per step ("final detection collate").  The reference has no
distributed code at all."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_range(n_samples: int, rank: int, world: int):
    """samples [lo, hi) owned by `rank` (contiguous, remainder to the first ranks)."""
    base, rem = divmod(n_samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(det: torch.Tensor, ndet: torch.Tensor) -> torch.Tensor:
    """det [B,A,6] + ndet [B] -> one fp32 buffer [B, A*6+1] (count in the last column)."""
    B = det.shape[0]
    return torch.cat([det.reshape(B, -1), ndet.to(det.dtype).view(B, 1)], dim=1).contiguous()


def unpack_detections(buf: torch.Tensor, A: int):
    B = buf.shape[0]
    det = buf[:, :A * 6].reshape(B, A, 6)
    ndet = buf[:, A * 6].round().to(torch.int32)
    return det, ndet


def all_gather_detections(det: torch.Tensor, ndet: torch.Tensor, group=None):
    """every rank receives the detections of all shards: det [world*B, A, 6], ndet [world*B]."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return det, ndet
    A = det.shape[1]
    mine = pack_detections(det, ndet)
    world = dist.get_world_size(group)
    out = torch.empty((world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return unpack_detections(out, A)


def detections_to_list(det: torch.Tensor, ndet: torch.Tensor) -> List[dict]:
    counts = ndet.tolist()
    out = []
    for b, n in enumerate(counts):
        d = det[b, :n]
        out.append(dict(boxes=d[:, :4], scores=d[:, 4], labels=d[:, 5].long()))
    return out
