"""GPU, world_size 2 on NCCL: the ONE collective of the path with real pipeline outputs in it.

Each rank runs the full pipeline on its contiguous shard and writes its results straight into OutputGather's send
buffer; after the all-gather every rank must hold, in global batch order, exactly what a single GPU computes for the
whole batch (pairs are independent, blocks_original.py has no cross-sample op).  Skipped on boxes with one GPU; the
host-side logic is covered on gloo by tests/test_parallel_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(B):
    g = torch.Generator().manual_seed(1234)
    return torch.rand(B, 6, 192, 256, generator=g) - 0.5


def _worker(rank, world, port, B, out_q):
    import torch.distributed as dist
    from demon_b200 import parallel, weights as W, _lib
    from demon_b200.networks_original import Session, DemonPipeline
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    r, local, w = parallel.init_from_env(backend="nccl")
    dev = torch.device("cuda", local)
    begin, end = parallel.shard_range(B, rank, world)
    x = _inputs(B)
    sess = Session()
    sess.load_weights(W.synthetic_weights(0))
    pipe = DemonPipeline(sess, batch_size=end - begin, iterations=3)
    gather = parallel.OutputGather(end - begin, world, device=dev)
    d, rot, tr = gather.local_buffers()
    pipe.forward(x[begin:end].to(dev), None, outputs={"predict_depth0": d, "predict_rotation": rot, "predict_translation": tr})
    gather()
    torch.cuda.synchronize()
    depth_all, motion_all = gather.gathered()
    res = {"rank": rank, "timeouts": _lib.load().demon_debug_tc_timeouts()}
    if rank == 0:   # the whole batch on this one GPU
        full = DemonPipeline(sess, batch_size=B, iterations=3, private_net=True).forward(x.to(dev), None)
        torch.cuda.synchronize()
        ref_d = full["predict_depth0"]
        ref_m = torch.cat([full["predict_rotation"], full["predict_translation"]], dim=1)
        res["depth_maxdiff"] = float((depth_all - ref_d).abs().max())
        res["motion_maxdiff"] = float((motion_all - ref_m).abs().max())
        res["depth_absmax"] = float(ref_d.abs().max())
        res["finite"] = bool(torch.isfinite(depth_all).all() and torch.isfinite(motion_all).all())
    # every rank holds the same gathered record
    chk = torch.stack([depth_all.double().sum(), motion_all.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["same_on_all_ranks"] = bool(torch.equal(lo, hi))
    out_q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_gather_holds_the_single_gpu_result_in_global_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    B = 4
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0 = [r for r in results if r["rank"] == 0][0]
    assert all(r["timeouts"] == 0 for r in results)
    assert all(r["same_on_all_ranks"] for r in results)
    assert r0["finite"]
    # same kernels, same per-sample arithmetic whatever the batch split: equal up to the last bit or two
    assert r0["depth_maxdiff"] <= 1e-6 * max(1.0, r0["depth_absmax"]), r0
    assert r0["motion_maxdiff"] <= 1e-6, r0
