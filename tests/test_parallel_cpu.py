"""N > 1 host logic on CPU: batch sharding and the one collective of the path (all_gather of padded
detections) with the gloo backend, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dagr_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, A = 3, 175
        g = torch.Generator().manual_seed(100 + rank)
        det = torch.rand((B, A, 6), generator=g)
        ndet = torch.tensor([5 + rank, 0, 175], dtype=torch.int32)
        for b in range(B):
            det[b, int(ndet[b]):] = 0
        gd, gn = parallel.all_gather_detections(det, ndet)
        q.put((rank, gd.clone(), gn.clone(), det, ndet))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for n in (8, 64, 7, 1):
        for world in (1, 2, 4, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_pack_unpack_roundtrip():
    det = torch.rand(4, 175, 6); nd = torch.tensor([0, 3, 175, 9], dtype=torch.int32)
    d2, n2 = parallel.unpack_detections(parallel.pack_detections(det, nd), 175)
    assert torch.equal(d2, det) and torch.equal(n2, nd)
    lst = parallel.detections_to_list(det, nd)
    assert [len(x["boxes"]) for x in lst] == [0, 3, 175, 9] and lst[1]["labels"].dtype == torch.long


def test_all_gather_detections_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_det = torch.cat([res[0][3], res[1][3]])
    want_n = torch.cat([res[0][4], res[1][4]])
    for rank, gd, gn, _, _ in res:          # every rank holds all detections, rank-major sample order
        assert torch.equal(gd, want_det) and torch.equal(gn, want_n)
