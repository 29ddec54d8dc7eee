"""CPU, world_size 2 on gloo: the host-side sharding / gather logic of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from demon_b200 import parallel


def test_shard_ranges_cover_the_batch():
    for B in (1, 2, 7, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            ranges = [parallel.shard_range(B, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_range(512, 3, 8) == (192, 256)       # BASELINE.json configs[3]: 64 pairs per GPU
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    r, _, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    begin, end = parallel.shard_range(B, rank, world)
    g = torch.Generator().manual_seed(1234)
    full_depth = torch.rand(B, 1, 6, 8, generator=g)
    full_motion = torch.rand(B, 6, generator=g)
    gather = parallel.OutputGather(end - begin, world, hw=(6, 8), device="cpu")
    # first through copies, then with the "pipeline" writing straight into the send buffer views
    gather(full_depth[begin:end], full_motion[begin:end, 0:3], full_motion[begin:end, 3:6])
    d_all, m_all = gather.gathered()
    ok = torch.equal(d_all, full_depth) and torch.equal(m_all, full_motion)
    d, r, t = gather.local_buffers()
    d.copy_(2 * full_depth[begin:end]); r.copy_(full_motion[begin:end, 0:3]); t.copy_(-full_motion[begin:end, 3:6])
    gather()
    d_all, m_all = gather.gathered()
    ok = ok and torch.equal(d_all, 2 * full_depth) and torch.equal(m_all[:, :3], full_motion[:, :3]) and torch.equal(m_all[:, 3:], -full_motion[:, 3:])
    ok = ok and gather.depth_all.shape == (world, end - begin, 1, 6, 8)
    try:
        gather(full_depth[:1], None, None)
        ok = False
    except ValueError:
        pass
    t = parallel.max_over_ranks(1.0 + rank, "cpu")
    out_q.put((rank, ok, t))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gather_reassembles_the_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results)
    assert all(r[2] == 2.0 for r in results)      # max over ranks
