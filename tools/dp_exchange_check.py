"""Multi-GPU check of the fused gradient exchange + optimiser kernel (csrc/dp_exchange.cu) on REAL symmetric memory, launched as
    torchrun --nproc-per-node N tools/dp_exchange_check.py
Every rank fills its gradient buffer with rank-dependent values, all ranks run barrier -> mfn_dp_exchange_adam -> barrier, and every rank
checks its master shard against torch (fp64 Adam on the summed gradient) and that all shadows agree with the masters of ALL ranks.  Both the NVSwitch multicast path (when the fabric has it) and the peer-pointer path are run and timed."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from mfnerf_b200 import dist as mdist
from mfnerf_b200._lib import call, ptr, stream_ptr


def main():
    rank, local, world = mdist.env_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 11448112 // (8 * world) * (8 * world)        # the Lego configuration's parameter count
    shard = n // world
    for use_mc in ("1", "0"):
        os.environ["MFN_DP_MULTICAST"] = use_mc
        sy = mdist.SymmetricBuffers(n, dev)
        if rank == 0:
            print(f"world {world}: multicast requested {use_mc} -> {'multimem (NVSwitch)' if sy.multicast else 'peer pointers'}", flush=True)
        if use_mc == "1" and not sy.multicast:
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        gm = torch.Generator(device=dev).manual_seed(7)                         # masters: the same on every rank
        p_full = torch.rand(n, device=dev, generator=gm) - 0.5
        p = p_full[rank * shard:(rank + 1) * shard].clone(); m = torch.zeros(shard, device=dev); v = torch.zeros(shard, device=dev)
        amp = torch.zeros(8, device=dev); call("mfn_amp_init", ptr(amp), 1024.0, 0, 0.9, 0.999, stream_ptr(dev))
        lr = torch.tensor([1e-2], device=dev); skip = torch.zeros(1, dtype=torch.int32, device=dev)
        sy.grads.copy_(torch.randn(n, device=dev, generator=g) * 30.0); sy.shadow.zero_(); sy.flag.zero_()
        mine = sy.grads.clone()
        allg = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allg, mine)
        want_g = torch.stack(allg).double().sum(0)[rank * shard:(rank + 1) * shard] / (1024.0 * world)
        torch.cuda.synchronize(); dist.barrier()
        times = []
        for it in range(4):
            if it > 0:                                                           # re-arm: same gradients again (only the timing of the later rounds counts)
                sy.grads.copy_(mine); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sy.barrier(0)
            e0.record()
            call("mfn_dp_exchange_adam", world, sy.grads_ptrs, sy.shadow_ptrs, sy.flag_ptrs, sy.grads_mc, sy.shadow_mc, ptr(p), ptr(m), ptr(v), rank * shard, shard,
                 ptr(lr), 0.9, 0.999, 1e-15, ptr(amp), ptr(skip), stream_ptr(dev))
            e1.record()
            sy.barrier(1)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
            if it == 0:
                mw = 0.1 * want_g; vw = 0.001 * want_g * want_g
                pw = p_full[rank * shard:(rank + 1) * shard].double() - 1e-2 * (mw / (1 - 0.9)) / (torch.sqrt(vw / (1 - 0.999)) + 1e-15)
                err_p = (p.double() - pw).abs().max().item(); err_m = (m.double() - mw).abs().max().item()
                zero = 0
                masters = [torch.empty_like(p) for _ in range(world)]
                dist.all_gather(masters, p)
                shadow_ok = torch.equal(sy.shadow, torch.cat(masters).half())
                ok = err_p < 5e-6 and err_m < 1e-6 and zero == 0 and shadow_ok and int(skip) == 0
                print(f"  rank {rank}: max|dp| {err_p:.2e} max|dm| {err_m:.2e} shadows equal masters of all ranks {shadow_ok} -> {'OK' if ok else 'FAIL'}", flush=True)
                assert ok
        t = torch.tensor([min(times[1:])], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            gb = 4.0 * n / world * (world - 1) / world + 2.0 * n / world * (world - 1)
            print(f"  exchange kernel {t.item():.1f} us (max over ranks, best of 3) for {n} parameters: "
                  f"{4 * n / 1e6:.1f} MB of gradients reduced, {2 * n / 1e6:.1f} MB of shadow broadcast", flush=True)
        del sy
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
