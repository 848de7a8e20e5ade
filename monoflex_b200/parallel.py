"""Multi-GPU plumbing of the hot path: images are independent, so inference shards the batch across one
process per GPU with NO data-path collective; the only cross-rank operations are the launch barrier and the
max-over-ranks reduction of the timed interval (bench.py). Training's gradient allreduce (row R13) is a later row.
Works on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """init_process_group from the torchrun environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE); no-op for world 1."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous shard of `n_items` for `rank` (InferenceSampler semantics, data/samplers/distributed_sampler.py:193-196):
    the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device="cpu"):
    """element-wise max of a list of floats over all ranks (timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
