"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: sharding + the single all-gather of key-point records."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hand3d_b200.distributed import (RECORD_FLOATS, gather_ragged_records, gather_records, pack_records, shard_range,
                                         unpack_records)
    g = torch.Generator().manual_seed(0)
    coord = torch.randn(total, 21, 3, generator=g)
    uv = torch.randint(0, 256, (total, 21, 2), generator=g, dtype=torch.int32)
    cen = torch.randn(total, 2, generator=g); sc = torch.rand(total, 1, generator=g)
    lo, hi = shard_range(total, rank, world)
    rec = pack_records(coord[lo:hi], uv[lo:hi], cen[lo:hi], sc[lo:hi])
    assert rec.shape == (hi - lo, RECORD_FLOATS) and rec.element_size() * RECORD_FLOATS == 432
    full = gather_ragged_records(rec, total) if total % world else gather_records(rec)
    out = unpack_records(full)
    ok = (torch.equal(out["keypoint_coord3d"], coord) and torch.equal(out["keypoints_uv"], uv) and
          torch.equal(out["center"], cen) and torch.equal(out["scale_crop"], sc))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_records_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_covers_everything():
    from hand3d_b200.distributed import shard_range
    for total in (1, 7, 8, 256, 513):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
