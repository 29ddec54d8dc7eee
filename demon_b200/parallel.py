"""Multi-GPU plumbing of the DeMoN inference path: image pairs are independent
(blocks_original.py has no cross-sample op; `predicted_scale` is per sample, blocks_original.py:281-283),
so a batch shards by contiguous ranges over one process per GPU and the only exchange is ONE all-gather
(one NCCL call per step) of the final depth / motion tensors.  The reference has no distributed code at all.

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
    """The single collective of the path: ONE all-gather per step of every rank's final predictions.

    A rank's record is one contiguous buffer [depth0 (b x H x W) | rotation (b x 3) | translation (b x 3)];
    `local_buffers()` hands out views into it so that the pipeline writes its results straight into the send buffer
    (no staging copies), and `__call__()` is one `all_gather_into_tensor` on the current stream -- capturable in a CUDA
    graph together with the pipeline that precedes it.  Shards must be equal (BASELINE.json configs[3]: 512 pairs over
    8 GPUs = 64 per rank); `shard_range` can produce ragged shards for other uses, this class refuses them.
    Results: `depth_all` [world, b, 1, H, W], `rotation_all` / `translation_all` [world, b, 3] (views, rank-major = the
    global batch order of `shard_range`); `gathered()` returns them reshaped to the global batch (copies)."""

    def __init__(self, shard_batch, world_size, hw=(192, 256), device="cuda"):
        self.world = int(world_size)
        self.shard = int(shard_batch)
        self.hw = tuple(hw)
        b, n = self.shard, self.hw[0] * self.hw[1]
        self.record = b * (n + 6)
        self.local = torch.empty(self.record, dtype=torch.float32, device=device)
        self.flat_all = self.local if self.world == 1 else torch.empty(self.world * self.record, dtype=torch.float32, device=device)
        self.all = self.flat_all.view(self.world, self.record)
        self.depth_all = self.all[:, :b * n].unflatten(1, (b, 1) + self.hw)
        self.rotation_all = self.all[:, b * n:b * n + 3 * b].unflatten(1, (b, 3))
        self.translation_all = self.all[:, b * n + 3 * b:].unflatten(1, (b, 3))

    def local_buffers(self):
        """(depth0 [b,1,H,W], rotation [b,3], translation [b,3]): views into this rank's send buffer."""
        b, n = self.shard, self.hw[0] * self.hw[1]
        return (self.local[:b * n].view((b, 1) + self.hw), self.local[b * n:b * n + 3 * b].view(b, 3), self.local[b * n + 3 * b:].view(b, 3))

    def __call__(self, depth0=None, rotation=None, translation=None):
        """Gathers the local record.  Tensors that are not the `local_buffers()` views are copied in first."""
        d, r, t = self.local_buffers()
        for src, dst in ((depth0, d), (rotation, r), (translation, t)):
            if src is not None and src.data_ptr() != dst.data_ptr():
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError("OutputGather: shard of shape %s, expected %s (shards must be equal)" % (tuple(src.shape), tuple(dst.shape)))
                dst.copy_(src)
        if self.world > 1:
            dist.all_gather_into_tensor(self.flat_all, self.local)
        return self.depth_all, self.rotation_all, self.translation_all

    def gathered(self):
        """(depth [world*b,1,H,W], motion [world*b,6]) of the last call, in global batch order (copies)."""
        wb = self.world * self.shard
        motion = torch.cat([self.rotation_all.reshape(wb, 3), self.translation_all.reshape(wb, 3)], dim=1)
        return self.depth_all.reshape((wb, 1) + self.hw), motion


def max_over_ranks(value, device):
    """Timing rule of the bench contract: the step time of a multi-GPU run is the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
