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


class SymmetricBuffers:
    """[fp32 gradients | fp16 shadow parameters | int32 overflow flag] of one rank in ONE symmetric-memory allocation
    (torch.distributed._symmetric_memory: the same virtual layout on every rank, every rank's copy peer-mapped over NVLink, and one
    NVSwitch multicast address for the whole allocation where the fabric supports it).  Gives csrc/dp_exchange.cu its pointer tables and the
    device-side barriers that bracket the exchange kernel."""

    def __init__(self, n_params, device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass                                    # newer torch: rendezvous() enables it by itself
        world = dist.get_world_size(group)
        al = lambda b: (b + 255) // 256 * 256
        off_g, off_s = 0, al(4 * n_params)
        off_f = off_s + al(2 * n_params)
        total = off_f + 256
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.grads = self.buf[off_g:off_g + 4 * n_params].view(torch.float32)
        self.shadow = self.buf[off_s:off_s + 2 * n_params].view(torch.float16)
        self.flag = self.buf[off_f:off_f + 4].view(torch.int32)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        arr = lambda off: (ctypes.c_uint64 * world)(*[b + off for b in base])
        self.grads_ptrs, self.shadow_ptrs, self.flag_ptrs = arr(off_g), arr(off_s), arr(off_f)
        # NVSwitch multicast (multimem.ld_reduce / multimem.st) is opt-in (MFN_DP_MULTICAST=1): measured on B200 with 11.4 M parameters the
        # in-fabric reduction costs ~175-200 us whatever the number of ranks, the peer-load path 93 us (2 GPUs), 114 us (4), 183 us (8) --
        # and at 8 GPUs, where the two kernels tie, the step with peer loads is still the shorter one (0.485 vs 0.495 ms).
        want = os.environ.get("MFN_DP_MULTICAST", "0")
        use_mc = getattr(self.hdl, "has_multicast_support", False) and want == "1"
        mc = int(self.hdl.multicast_ptr) if use_mc else 0
        self.grads_mc, self.shadow_mc = (mc + off_g, mc + off_s) if mc else (0, 0)
        self.multicast = bool(mc)

    def barrier(self, channel):
        """device-side barrier of all ranks on the current stream (signal pads of the symmetric allocation)"""
        self.hdl.barrier(channel=channel)
