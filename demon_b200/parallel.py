"""Multi-GPU plumbing of the DeMoN inference path: image pairs are independent
(blocks_original.py has no cross-sample op; `predicted_scale` is per sample, blocks_original.py:281-283),
so a batch shards by contiguous ranges over one process per GPU and the only exchange is ONE all-gather
of the final depth / motion tensors.  The reference has no distributed code at all.

Works on the `nccl` backend (CUDA tensors, the product path) and on `gloo` (CPU tensors, used by the
world_size-2 tests of this host logic)."""
import os

import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world_size):
    """Contiguous, balanced shard [begin, end) of `global_batch` pairs for `rank` (first ranks get the remainder)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(global_batch, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """One process per GPU as launched by torch.distributed.run; returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


class OutputGather:
    """The single collective of the path: all-gather of depth0 [b,1,H,W] and motion (rotation|translation)
    [b,6] of equal-sized shards into preallocated global buffers, launched on the current stream."""

    def __init__(self, shard_batch, world_size, hw=(192, 256), device="cuda"):
        self.world = world_size
        self.shard = shard_batch
        self.depth_all = torch.empty((world_size * shard_batch, 1) + tuple(hw), dtype=torch.float32, device=device)
        self.motion = torch.empty((shard_batch, 6), dtype=torch.float32, device=device)
        self.motion_all = torch.empty((world_size * shard_batch, 6), dtype=torch.float32, device=device)

    def __call__(self, depth0, rotation, translation):
        self.motion[:, 0:3].copy_(rotation)
        self.motion[:, 3:6].copy_(translation)
        if self.world == 1:
            self.depth_all.copy_(depth0)
            self.motion_all.copy_(self.motion)
        else:
            dist.all_gather_into_tensor(self.depth_all, depth0.contiguous())
            dist.all_gather_into_tensor(self.motion_all, self.motion)
        return self.depth_all, self.motion_all


def max_over_ranks(value, device):
    """Timing rule of the bench contract: the step time of a multi-GPU run is the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
