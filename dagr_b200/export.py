"""Recover the reference's data layout from the engine's internal state (parity checks, API
compatibility -- NOT on the hot path; plain torch ops, host syncs allowed).

  * event level: node id = arrival index; `perm[p]` maps cell-major sorted position -> arrival index
  * coarse levels: consecutive node ids in ascending voxel id order (pooling.py:12-16,57) and the
    lexicographically sorted unique coarse edge list (pooling.py:58-64)
"""
from __future__ import annotations

import torch


def unsort_rows(x_sorted: torch.Tensor, perm: torch.Tensor, N: int) -> torch.Tensor:
    """rows in sorted-position order -> arrival order."""
    out = torch.empty_like(x_sorted[:N])
    out[perm[:N].long()] = x_sorted[:N]
    return out


def grid_nodes(gs, level, geom):
    """dict(valid, ids, batch, pos_px [n,2] int, pos [n,2] float normalised, x [n,C], cell [n])."""
    valid = gs.cnt[:gs.cells] > 0
    ids = torch.cumsum(valid.long(), 0) - 1
    cell = torch.nonzero(valid).flatten()
    per = level.nx * level.ny
    px = gs.pxy[cell, 0].long()
    py = gs.pxy[cell, 1].long()
    pos = torch.stack([geom.d_posxr[px], geom.d_posyr[py]], dim=1)
    return dict(valid=valid, ids=ids, cell=cell, batch=cell // per, pos_px=torch.stack([px, py], 1), pos=pos,
                x=None if gs.x is None else gs.x[cell], tmean=gs.tmean[cell])


def grid_edges(gs, level):
    """edge_index int64 [2,E] over consecutive ids, sorted by (src, dst) like torch.unique(dim=-1)."""
    valid = gs.cnt[:gs.cells] > 0
    ids = torch.cumsum(valid.long(), 0) - 1
    nx, ny = level.nx, level.ny
    per = nx * ny
    cell = torch.arange(gs.cells, device=gs.cnt.device)
    rem = cell % per
    cy, cx = rem // nx, rem % nx
    mask = gs.mask[:gs.cells].long()
    src_l, dst_l = [], []
    for bit in range(9):
        if bit == 4:
            continue
        dcx, dcy = bit % 3 - 1, bit // 3 - 1
        has = ((mask >> bit) & 1).bool() & valid
        sx, sy = cx + dcx, cy + dcy
        ok = has & (sx >= 0) & (sy >= 0) & (sx < nx) & (sy < ny)
        src = cell + dcx + dcy * nx
        ok = ok & valid[src.clamp(0, gs.cells - 1)]
        src_l.append(ids[src[ok]])
        dst_l.append(ids[cell[ok]])
    src = torch.cat(src_l)
    dst = torch.cat(dst_l)
    n = int(valid.sum())
    order = torch.argsort(src * max(n, 1) + dst)
    return torch.stack([src[order], dst[order]])
