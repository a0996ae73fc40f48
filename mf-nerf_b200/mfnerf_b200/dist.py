"""Host-side logic of the data-parallel path: one process per GPU, rays sharded across ranks, ONE
all-reduce of the flat gradient per training step, image row tiles at test time (no collective).

Nothing here touches CUDA directly, so the same code runs under `gloo` on CPU tensors (tests/test_multi_cpu.py) and under
`nccl` on the GPUs (engine.py, bench.py)."""
import os
import time

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_seed(base_seed, rank):
    """every rank samples its OWN rays (the reference's per-rank dataloader sampling, datasets/base.py:22-44)"""
    return int(base_seed) + 1000003 * int(rank)


def model_seed(base_seed):
    """...but the SAME initial parameters and the same density-grid cell sampling on every rank, so the replicas stay identical"""
    return int(base_seed)


def allreduce_gradients(flat_grads, world_size, group=None, overflow_flag=None):
    """sum the flat gradient [xyz_encoder.params | rgb_net.params] over ranks, in place.  The division by world_size (DDP's
    gradient averaging, train.py uses Lightning DDP) is folded into the optimiser's grad_scale -- see grad_scale().
    `overflow_flag` (int32 tensor, 1 = an fp16 gradient overflowed on this rank) is max-reduced so that every replica takes
    the same skip-or-step decision (GradScaler semantics under DDP)."""
    if world_size > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        if overflow_flag is not None:
            dist.all_reduce(overflow_flag, op=dist.ReduceOp.MAX, group=group)
    return flat_grads


def grad_scale(loss_scale, world_size):
    """factor the fused Adam applies to the summed, loss-scaled gradient"""
    return 1.0 / (float(loss_scale) * int(world_size))


def tile_rows(height, rank, world_size):
    """[row0, row1) of the image rows rendered by `rank` (contiguous row tiles; remainder rows go to the first ranks)"""
    base, rem = divmod(int(height), int(world_size))
    row0 = rank * base + min(rank, rem)
    return row0, row0 + base + (1 if rank < rem else 0)


def max_over_ranks(value, device, world_size, group=None):
    """max of a python float over ranks (bench.py: the step time of a job is its slowest rank's)"""
    if world_size == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, device, world_size, group=None):
    if world_size == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def agreed_warmup(step, world_size, device, seconds, chunk=128, max_chunks=64, group=None, clock=time.perf_counter, sync=None):
    """Untimed host warm-up whose LENGTH is the same on every rank: `step()` runs in chunks of `chunk`; after each chunk rank 0's
    clock decides whether `seconds` have passed and the decision is broadcast.  Every training step enqueues collectives when
    world_size > 1, so a per-rank wall-clock loop would let the ranks issue different numbers of them and dead-lock (bench.py,
    round 1).  Returns the number of steps run."""
    t0, n = clock(), 0
    for _ in range(int(max_chunks)):
        for _ in range(int(chunk)):
            step()
        n += int(chunk)
        if sync is not None:
            sync()
        stop = (clock() - t0) >= seconds
        if world_size > 1:
            t = torch.tensor([1 if stop else 0], device=device, dtype=torch.int32)
            dist.broadcast(t, src=0, group=group)
            stop = bool(int(t.item()))
        if stop:
            break
    return n
