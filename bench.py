#!/usr/bin/env python
"""bench.py -- training rays/s (and samples/s, 800x800 render FPS) of the MF-NeRF hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference [...]                          the reference path on the host cores (port, see oracle/cpu_step.py)
    torchrun --nproc-per-node N bench.py --gpus N ...               one rank per GPU, rays sharded, NCCL all-reduce of the gradients

Workload (BASELINE.json configs[1]): Lego-shaped synthetic scene, 800x800 pinhole cameras on a hemisphere of radius 1.5, 8192 rays
per step per GPU, scale 0.5 (one cascade, dt = sqrt(3)/1024), hash grid L16 F2 T2^19, 64x1 sigma net, 64x2 rgb net, random-init weights,
analytic box-scene colours as targets.  One "step" = the reference's training_step (train.py:164-190): [update_density_grid every 16
steps] -> AABB -> march -> field -> composite -> loss -> backward -> [all-reduce] -> Adam.  The occupancy grid starts from the procedural
Lego voxelisation, i.e. the post-warm-up regime, and then evolves through the real update rule.

value  = rays/s over K steps with every ray batch already resident in HBM (a pool of distinct batches, cycled);
e2e    = the same K steps fed from pinned HOST memory (H2D of rays+targets every step, D2H of the loss every step);
roofline / breakdown = per-kernel CUDA-event times recorded inside the replayed CUDA graph in a separate pass of the same steps.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

R_PER_GPU = 8192
POOL = 64                      # distinct ray batches cycled through (64 x 8192 rays)
METRIC = "train_rays_per_sec"
UNIT = "rays/s"
WORKLOAD = "lego_synthetic_train_8192rays_scale0.5_hashL16F2T19"
RENDER_MIN_CHUNK = 8           # lower bound of the per-iteration sample count of the test-time renderer (see engine.render)

# --config: the workload.  "lego" is BASELINE.json configs[1], the one the metric is quoted on and the default; the others are the
# remaining single-GPU-sized shapes of BASELINE.json (configs 3-5) and MF-NeRF's own script configuration, benchable on 1..8 GPUs.
CONFIGS = {
    "lego": dict(workload=WORKLOAD, rays=8192, engine=dict(scale=0.5, log2_T=19)),
    # benchmarking/benchmark_synthetic_nerf_mf.sh: --batch_size 16384 --lr 2e-2 --T 20 --grid MixedFeature --N_tables 8 --rgb_channels 128 --rgb_layers 2
    "mf_synthetic": dict(workload="lego_synthetic_train_16384rays_scale0.5_mixedfeatureK8_L16F2T20_rgb128x2", rays=16384,
                         engine=dict(scale=0.5, log2_T=20, grid="MixedFeature", n_tables=8, rgb_channels=128, rgb_layers=2, lr=2e-2)),
    # benchmarking/benchmark_synthetic_nerf_hash.sh: the fork's hash-grid control (T 20, 64x2 rgb net, 16384 rays)
    "hash_synthetic": dict(workload="lego_synthetic_train_16384rays_scale0.5_hashL16F2T20", rays=16384, engine=dict(scale=0.5, log2_T=20, lr=2e-2)),
    # BASELINE.json configs[3]: mipnerf360-shaped unbounded scene, scale 16 (6 cascades, exp_step_factor 1/256), T 2^21, distortion loss
    "unbounded_T21": dict(workload="mipnerf360_shaped_train_8192rays_scale16_6cascades_hashL16F2T21_distortion", rays=8192,
                          engine=dict(scale=16.0, log2_T=21, distortion_w=1e-3)),
    # BASELINE.json configs[4]: forward-facing (LLFF) shape with distortion loss; --rays / --log2-T sweep the batch (2^13..2^20) and table (2^19..2^22)
    "llff_distortion": dict(workload="llff_shaped_train_scale4_4cascades_hashL16F2_distortion", rays=8192, engine=dict(scale=4.0, log2_T=19, distortion_w=1e-3)),
    # benchmarking/benchmark_mipnerf360_mf.sh shape: unbounded scene on the MixedFeature grid, T 22, 128x2 rgb net
    "mf_unbounded_T22": dict(workload="mipnerf360_shaped_train_8192rays_scale16_mixedfeatureK8_L16F2T22_rgb128x2_distortion", rays=8192,
                             engine=dict(scale=16.0, log2_T=22, grid="MixedFeature", n_tables=8, rgb_channels=128, rgb_layers=2, distortion_w=1e-3)),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--config", default="lego", choices=sorted(CONFIGS), help="workload (default: BASELINE.json configs[1])")
    ap.add_argument("--rays", type=int, default=0, help="rays per step per GPU (default: the config's)")
    ap.add_argument("--log2-T", type=int, default=0, help="log2 of the hash-table size (default: the config's)")
    ap.add_argument("--no-python-layer", action="store_true", help="skip the gpu_reference / frozen_api legs (the reference's python layer on the GPU)")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sust=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ synthetic data
def make_pool(n_batches, rays, seed):
    """-> float32 numpy (n_batches, 3, rays, 3): [rays_o, rays_d, target] per batch"""
    import numpy as np
    from mfnerf_b200 import synthetic as syn
    o, d, _, _ = syn.random_rays(n_batches * rays, seed=seed)
    tgt = syn.analytic_render(o, d).numpy()
    pool = np.stack([o.reshape(n_batches, rays, 3), d.reshape(n_batches, rays, 3), tgt.reshape(n_batches, rays, 3)], 1)
    return np.ascontiguousarray(pool, dtype=np.float32)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons with NVML while the timed region runs"""

    def __init__(self, torch_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80, "sw_power_cap": 0x4}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ reference arm (host cores)
def cpu_reference(steps, warmup, rays=R_PER_GPU, budget_s=25.0, warm_budget_s=None):
    """the reference training step as a CPU port (oracle/cpu_step.py) on all host cores -> (rays/s, dict).
    Same schedule as the GPU arm: steps count from 1 from the procedural Lego occupancy grid, update_density_grid every 16 steps (all
    cells while step < 256), `warmup` untimed steps first -- so that both arms are timed in the same occupancy regime (round 1 timed
    the CPU arm on the untouched grid: 25 samples/ray against the GPU arm's 58).  The warm-up is cut at `warm_budget_s` of wall time
    (reported: warmup_steps_run, same_schedule)."""
    import numpy as np
    from mfnerf_b200 import synthetic as syn
    from oracle.cpu_step import CpuTrainer
    if warm_budget_s is None:
        warm_budget_s = float(os.environ.get("MFN_REF_WARM_BUDGET_S", "75"))
    budget_s = float(os.environ.get("MFN_REF_BUDGET_S", budget_s))
    tr = CpuTrainer(scale=0.5, log2_T=19, threads=os.cpu_count())
    tr.set_density_grid(syn.lego_density_grid(0.5, 1))
    pool = make_pool(2, rays, seed=101)
    step, t0 = 1, time.perf_counter()
    for _ in range(max(1, warmup)):
        tr.training_step(step, pool[step % 2, 0], pool[step % 2, 1], pool[step % 2, 2]); step += 1
        if time.perf_counter() - t0 > warm_budget_s:
            break
    warm_run = step - 1
    t0 = time.perf_counter()
    done = samples = 0
    for _ in range(steps):
        _, n = tr.training_step(step, pool[step % 2, 0], pool[step % 2, 1], pool[step % 2, 2]); step += 1
        done += 1; samples += n
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * rays / dt, dict(steps_run=done, seconds=dt, samples_per_ray=samples / max(1, done * rays), cores=tr.threads, warmup_steps_run=warm_run,
                                  same_schedule=warm_run >= warmup,
                                  sample=f"{done} full training steps of {rays} rays (occupancy update every 16 steps, AABB, march, field fwd/bwd, composite fw/bw, "
                                         f"loss, Adam) of the same Lego-shaped workload after {warm_run} untimed warm-up steps, {dt:.1f} s of CPU work")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 24)
    v, info = cpu_reference(steps, max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": info["steps_run"], "warmup": info["warmup_steps_run"],
        "ms_per_step": 1e3 * info["seconds"] / max(1, info["steps_run"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": R_PER_GPU, "density_grid_update_every": 16, "same_schedule_as_gpu_arm": info["same_schedule"],
                   "note": "vren/tcnn have no CPU kernels: C restatement of vren + torch transliteration of tcnn (oracle/), all host threads"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_ray": info["samples_per_ray"], "samples_per_sec": v * info["samples_per_ray"],
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ the reference's python layer on the GPU
def python_layer_arm(kind, dev, pool_dev, steps, warmup):
    """The reference's UNMODIFIED python layer (models/rendering.py, custom_functions.py, networks.py, losses.py, staged under
    oracle/_ref/refpy by oracle/build_ref_vren.sh) driving one training_step after the other (train.py:164-190) on the same ray
    batches as our arm, under autocast + GradScaler (Lightning precision=16) and Adam(eps=1e-15):
      kind = "gpu_reference": on the reference's OWN vren kernels recompiled for sm_100 (oracle/_ref/vren_ref*.so) + a torch
             transliteration of the un-vendored tiny-cuda-nn fork (oracle/tcnn_torch.py) -- BASELINE.md's "GPU-ref A+B";
      kind = "frozen_api"   : on this repo's drop-in modules (vren / tinycudann / torch_scatter over the C ABI) -- what a user
             gets by running the reference's train.py unchanged.
    Timed with CUDA events around `steps` steps (host syncs of the reference flow included), outside our arm's timed regions."""
    import contextlib
    import glob
    import importlib.util
    import types
    import torch
    refpy = os.path.join(ROOT, "oracle", "_ref", "refpy")
    if not os.path.isdir(os.path.join(refpy, "models")):
        return {"unavailable": "oracle/_ref/refpy not staged (oracle/build_ref_vren.sh runs where /root/reference exists)"}
    names = ("vren", "tinycudann", "torch_scatter", "losses", "models")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in names or k.startswith("models.")}
    try:
        if kind == "gpu_reference":
            so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "vren_ref*.so"))
            if not so:
                return {"unavailable": "oracle/_ref/vren_ref*.so not built"}
            spec = importlib.util.spec_from_file_location("vren_ref", so[0])
            vren_mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(vren_mod)
            from oracle import tcnn_torch
            scatter = types.ModuleType("torch_scatter")
            scatter.segment_csr = lambda src, indptr: torch.segment_reduce(src, "sum", offsets=indptr, axis=0)
            sys.modules.update(vren=vren_mod, tinycudann=tcnn_torch, torch_scatter=scatter)
            label = "reference vren kernels recompiled for sm_100 + torch transliteration of tiny-cuda-nn (tcnn itself is not in the tree)"
        else:
            import vren as vren_mod  # noqa: F401  (this repo's drop-ins, mf-nerf_b200/ on sys.path)
            import tinycudann  # noqa: F401
            import torch_scatter  # noqa: F401
            label = "reference python layer, unchanged, on this repo's vren / tinycudann / torch_scatter drop-ins"
        sys.path.insert(0, refpy)
        try:
            with contextlib.redirect_stdout(sys.stderr):
                from models import networks, rendering
                import losses
        finally:
            sys.path.remove(refpy)
        from mfnerf_b200 import synthetic as syn
        hp = types.SimpleNamespace(L=16, F=2, T=19, N_min=16, N_max=2048, N_tables=1, grid="Hash", rgb_channels=64, rgb_layers=2)
        with contextlib.redirect_stdout(sys.stderr):
            model = networks.NGP(scale=0.5, hparams=hp, rgb_act="Sigmoid").to(dev)
        G = model.grid_size
        model.register_buffer("density_grid", torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev).reshape(model.cascades, G ** 3).contiguous())   # train.py:78-81
        ar = torch.arange(G, dtype=torch.int32, device=dev)
        model.register_buffer("grid_coords", torch.stack(torch.meshgrid(ar, ar, ar, indexing="ij"), -1).reshape(-1, 3).contiguous())
        vren_mod.packbits(model.density_grid.reshape(-1).contiguous(), 0.5, model.density_bitfield)
        loss_fn = losses.NeRFLoss(lambda_distortion=0)                                  # opt.py:25 default
        opt = torch.optim.Adam([p for p in model.parameters() if p.numel() > 0], lr=1e-2, eps=1e-15)   # train.py:136 (apex FusedAdam is not installed)
        scaler = torch.amp.GradScaler("cuda")
        state = {"step": 1, "samples": 0}

        def step():
            s = state["step"]
            b = pool_dev[s % pool_dev.shape[0]]
            with torch.autocast("cuda", dtype=torch.float16):                            # Lightning precision=16 wraps the whole training_step
                if s % 16 == 0:                                                         # train.py:165-168
                    model.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=s < 256, erode=False)
                res = rendering.render(model, b[0], b[1], test_time=False, random_bg=False)
                ld = loss_fn(res, {"rgb": b[2]})
                loss = sum(v.mean() for v in ld.values())
            opt.zero_grad(set_to_none=True)
            scaler.scale(loss).backward()
            scaler.step(opt); scaler.update()
            state["samples"] += int(res["rm_samples"])
            state["step"] = s + 1
            return loss

        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        state["samples"] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        R = pool_dev.shape[2]
        out = {"value": steps * R / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "kind": label,
               "samples_per_ray": state["samples"] / (steps * R), "samples_per_sec": state["samples"] / (ms * 1e-3), "final_loss": float(loss),
               "optimizer": "torch.optim.Adam(eps=1e-15) + torch.amp.GradScaler (apex FusedAdam / Lightning are not installed)"}
        del model, opt
        torch.cuda.empty_cache()
        return out
    except Exception as e:      # a baseline leg must never take the main line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    finally:
        for k in [k for k in list(sys.modules) if k in names or k.startswith("models.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


# ------------------------------------------------------------------------------------------------ our arm
# algorithmic bytes per sample / per ray / per parameter, or FLOPs per sample, of every profiled kernel
def algorithmic(name, n_samples, n_rays, n_params, rgb_flops):
    tab = {
        "grid_encode_fwd": ("hbm", 588.0 * n_samples), "grid_encode_bwd": ("hbm", 588.0 * n_samples),
        "march_count": ("hbm", 60.0 * n_rays + 8.0 * n_samples), "march_write": ("hbm", 24.0 * n_rays + 40.0 * n_samples),
        "composite_train_fw": ("hbm", 28.0 * n_samples + 52.0 * n_rays), "composite_train_bw": ("hbm", 48.0 * n_samples + 64.0 * n_rays),
        "composite_loss_train": ("hbm", 64.0 * n_samples + 128.0 * n_rays),     # fw 28 B + second sweep 20 B read, 16 B written per sample
        "adam": ("hbm", 30.0 * n_params),
        "mlp_sigma_fwd": ("tensor", 6144.0 * n_samples), "mlp_rgb_fwd": ("tensor", rgb_flops * n_samples),
        "mlp_sigma_bwd": ("tensor", 2 * 6144.0 * n_samples), "mlp_rgb_bwd": ("tensor", 2 * rgb_flops * n_samples),
        "field_fwd": ("hbm", 588.0 * n_samples), "field_bwd": ("hbm", 588.0 * n_samples),
    }
    return tab.get(name)


PROFILED = ["march_count", "march_scan", "march_write", "grid_encode_fwd", "mlp_sigma_fwd", "sh_sigma", "mlp_rgb_fwd", "field_fwd", "composite_train_fw",
            "nerf_loss", "composite_train_bw", "composite_loss_train", "prep_dout2", "mlp_rgb_bwd", "merge_dh", "mlp_sigma_bwd", "grid_encode_bwd", "field_bwd", "adam"]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mfnerf_b200 import _lib
    from mfnerf_b200 import dist as mdist
    from mfnerf_b200 import synthetic as syn
    from mfnerf_b200.engine import NGPEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the host-core port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    conf = CONFIGS[args.config]
    ekw = dict(conf["engine"])
    if args.log2_T:
        ekw["log2_T"] = args.log2_T
    R = args.rays or conf["rays"]
    is_default = args.config == "lego" and R == R_PER_GPU and not args.log2_T
    workload = conf["workload"] + ("" if is_default else f"_R{R}_T{ekw['log2_T']}")
    sampler = ClockSampler(local)      # NVML is initialised here, long before the timed regions (its start-up disturbs kernel launches for a while)

    sample_capacity = None if R <= 16384 else R * 256      # large batches: 256 samples per ray on average instead of the worst case 1024
    eng = NGPEngine(n_rays=R, device=dev, world_size=world, seed=1337, sample_capacity=sample_capacity, **ekw)
    eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(ekw["scale"], eng.cascades)).to(dev))
    eng.repack_bitfield(0.5)
    pool_n = max(4, min(POOL, (1 << 23) // R))     # distinct ray batches cycled through (64 at the default batch size)
    pool_host = torch.from_numpy(make_pool(pool_n, R, seed=1000 + rank)).pin_memory()
    pool_dev = pool_host.to(dev)
    loss_host = torch.zeros(K + W + 8, 3).pin_memory()
    torch.cuda.synchronize(dev)

    # steps start at 1: step 0 would be the reference's first update_density_grid from an untrained network (warm-up regime)
    state = {"step": 1}

    def step_from(pool, i_host_loss=None):
        s = state["step"]
        eng.train_step_packed(pool[s % pool_n], global_step=s)
        if i_host_loss is not None:
            eng.loss_to_host(loss_host[i_host_loss])
        state["step"] = s + 1

    # ---- warm-up (eager first so that lazy initialisation happens outside capture), then capture
    for _ in range(3):
        step_from(pool_dev)
    if not args.no_graph:
        eng.capture()
    launches_per_fb = eng.launches_per_forward_backward
    for _ in range(W):
        step_from(pool_dev)
    # first use of the post-warm-up occupancy update (networks.py:170-197 cell sampling: cumsum / searchsorted / randint ...) loads a
    # dozen torch kernels lazily, ~0.1 s of host time: do it here, not inside the timed region (steps >= 256 take that path)
    eng.update_density_grid(warmup=False)
    torch.cuda.synchronize(dev)
    snap, snap_step = eng.snapshot(), state["step"]       # every timed region below restarts from this training state
    # host warm-up: on a fresh box the first second of stepping is host-bound (libcuda / interpreter pages still faulting in: ~0.4 ms
    # of launch work per step instead of ~0.13 ms), which would be charged to the first timed region.  Keep stepping, untimed, for
    # 1.5 s of wall time; every timed region restarts from the snapshot taken BEFORE it (same workload as without it).
    # Every step enqueues collectives when world > 1, so the NUMBER of steps must be identical on every rank: the wall clock only
    # decides between chunks, and the decision is rank 0's, broadcast to the others (a per-rank clock would let ranks issue different
    # numbers of reduce-scatters and dead-lock the job -- round 1's SCALE failure).
    extra_warmup = mdist.agreed_warmup(lambda: step_from(pool_dev), world, dev, seconds=1.5, chunk=128, max_chunks=64)
    torch.cuda.synchronize(dev)

    def rewind():
        eng.restore(snap); state["step"] = snap_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(pool, with_loss_readback):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rewind()
        eng.reset_samples_marched()
        l0 = _lib.lib.mfn_launch_count(); g0 = eng.graph_replays
        barrier()
        ev0.record()
        t_cpu = time.perf_counter()
        for i in range(K):
            step_from(pool, i if with_loss_readback else None)
        t_cpu = time.perf_counter() - t_cpu
        eng.flush()        # the timing stream waits for the last step's backward pass, optimiser and loss read-back on the engine's streams
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if os.environ.get("MFN_BENCH_VERBOSE") and rank == 0:
            print(f"[timed] gpu {ms / K:.4f} ms/step, cpu launch loop {1e3 * t_cpu / K:.4f} ms/step", file=sys.stderr, flush=True)
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        launches = (_lib.lib.mfn_launch_count() - l0) + (eng.graph_replays - g0) * launches_per_fb
        samples = float(eng.samples_marched().item())
        if world > 1:
            t = torch.tensor([samples], device=dev, dtype=torch.float64); dist.all_reduce(t); samples = float(t.item())
        return ms, launches, samples

    # ---- device-resident run (value)
    if not os.environ.get("MFN_BENCH_NO_SAMPLER"):
        sampler.start()
    if os.environ.get("MFN_BENCH_SWAP"):
        for _ in range(int(os.environ["MFN_BENCH_SWAP"])):
            timed(pool_dev, False)
    ms_dev, launches, samples_dev = timed(pool_dev, False)
    clocks = sampler.stop()
    amp = eng.amp_state()                # loss scale and the number of steps skipped for fp16 overflow since the engine was built
    # ---- end-to-end run: rays + targets from pinned host memory every step, loss read back every step
    if not os.environ.get("MFN_BENCH_SWAP"):
        ms_e2e, _, _ = timed(pool_host, True)
    rays_total = float(K) * R * world
    value = rays_total / (ms_dev * 1e-3)
    e2e = rays_total / (ms_e2e * 1e-3)
    final_loss = loss_host[K - 1].tolist()

    # ---- per-kernel breakdown: CUDA events recorded inside the (re-captured) graph, one synchronised replay per step
    breakdown, roofline = {}, None
    pk = peaks()
    if rank == 0:
        evs = {}
        for name in PROFILED:
            a, b = _lib.lib.mfn_event_create(), _lib.lib.mfn_event_create()
            evs[name] = (a, b)
            _lib.check(_lib.lib.mfn_profile_set(name.encode(), a, b), "mfn_profile_set")
    if rank == 0 and not args.no_graph:
        eng.capture()
    n_prof = min(K, 48)
    rewind()
    acc = {n: [] for n in PROFILED}
    smp = []
    for i in range(n_prof):
        step_from(pool_dev)
        if rank == 0:
            torch.cuda.synchronize(dev)
            smp.append(int(eng.counter[0].item()))
            ms = ctypes.c_float()
            for name, (a, b) in evs.items():
                if _lib.lib.mfn_event_elapsed_ms(a, b, ctypes.byref(ms)) == 0:
                    acc[name].append(ms.value)
    if rank == 0:
        _lib.lib.mfn_profile_set(b"", None, None)
        if not args.no_graph:
            eng.capture()
        mean_s = float(np.mean(smp)) if smp else 0.0
        rgb_flops = 2.0 * (32 * 64 + 64 * 64 + 64 * 3)
        for name in PROFILED:
            if acc[name]:
                breakdown[name] = round(float(np.mean(acc[name][2:] or acc[name])) * 1e3, 2)     # microseconds
        if breakdown:
            top = max(breakdown, key=breakdown.get)
            alg = algorithmic(top, mean_s, R, eng.n_params, rgb_flops)
            if alg:
                bound, work = alg
                dur = breakdown[top] * 1e-6
                if bound == "hbm":
                    ach, peak, unit = work / dur / 1e9, pk["hbm"], "GB/s"
                else:
                    ach, peak, unit = work / dur / 1e12, pk["tf_sust"], "TFLOP/s"
                traffic = None
                try:
                    tj = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
                    if top in tj:
                        traffic = tj[top]["dram_bytes_per_sample"] * mean_s
                except Exception:
                    pass
                roofline = {"kernel": top, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": traffic,
                            "peak_source": pk["src"], "launch_us": breakdown[top], "algorithmic_per_launch": work, "samples_per_launch": mean_s}

    # ---- the two table kernels against the on-chip rooflines they are actually bound by ("for T <= 2^20 report achieved L2
    #      GB/s against a builder-measured random-32-B-sector L2 peak"): tools/l2_bench.cu -> profiles/l2_peaks_r02.json.  Algorithmic
    #      sector requests per sample: 16 levels x 4 (y,z) corner pairs (the x-neighbours of a pair share a sector 7 times out of 8).
    l2_roofline = None
    if rank == 0 and breakdown:
        try:
            lp = json.load(open(os.path.join(ROOT, "profiles", "l2_peaks_r02.json")))["_summary"]
            sectors = 64.0 * mean_s
            l2_roofline = {"sectors_per_sample": 64, "samples_per_launch": mean_s, "peak_source": "profiles/l2_peaks_r02.json (tools/l2_bench.cu, measured on B200)"}
            for kname, key, label in (("field_fwd", "l1_divergent_sector_peak_Gsectors_per_s", "gather"), ("grid_encode_bwd", "l2_red_v2_f32_peak_Gsectors_per_s", "red")):
                if kname in breakdown:
                    ach = sectors / (breakdown[kname] * 1e-6) / 1e9
                    l2_roofline[kname] = {"kind": label, "achieved_Gsectors_per_s": ach, "peak_Gsectors_per_s": lp[key], "frac": ach / lp[key], "launch_us": breakdown[kname]}
        except Exception:
            l2_roofline = None

    # ---- 800x800 test-time render (second half of the metric), on the trained state
    # ---- 800x800 test-time render (second half of the metric), on the trained state.  N > 1: every rank renders its tile of image rows
    #      (mfnerf_b200.dist.tile_rows, no collective); a frame's time is its slowest rank's
    render = None
    if not args.no_render:
        pose = syn.camera_poses(4, seed=7)
        row0, row1 = mdist.tile_rows(800, rank, world)
        frames, out = [], None
        for k in range(4):
            o, d = syn.image_rays(pose[k])
            o = torch.from_numpy(o.reshape(800, 800, 3)[row0:row1].reshape(-1, 3).copy()).to(dev)
            d = torch.from_numpy(d.reshape(800, 800, 3)[row0:row1].reshape(-1, 3).copy()).to(dev)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.render(o, d, min_chunk=RENDER_MIN_CHUNK); e1.record(); torch.cuda.synchronize(dev)
            frames.append(mdist.max_over_ranks(e0.elapsed_time(e1), dev, world))
        ms_frame = float(np.mean(frames[1:]))
        tot = mdist.sum_over_ranks(float(out["total_samples"]), dev, world)
        rows = mdist.sum_over_ranks(float(out.get("field_rows", 0)), dev, world)
        render = {"fps_800x800": 1e3 / ms_frame, "ms_per_frame": ms_frame, "samples_per_ray": tot / (800 * 800), "iterations": out.get("iterations"),
                  "field_rows_per_ray": rows / (800 * 800), "tiles": f"{world} row tile(s), no collective",
                  "schedule": f"N_samples = clamp(N_rays / N_alive, {RENDER_MIN_CHUNK}, 64) per iteration (reference: lower bound 1; same image, fewer iterations)"}

    # ---- the reference's python layer on the same GPU, same batches (rank 0 of a 1-GPU run only): GPU-ref A+B and the frozen API
    gpu_ref = frozen = None
    if rank == 0 and world == 1 and not args.no_python_layer and is_default:
        eng.flush(); torch.cuda.synchronize(dev)
        n_ref, w_ref = min(K, 64), min(W, 64)
        gpu_ref = python_layer_arm("gpu_reference", dev, pool_dev, n_ref, w_ref)
        frozen = python_layer_arm("frozen_api", dev, pool_dev, n_ref, w_ref)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and is_default:
        v, info = cpu_reference(12, W)
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"], "samples_per_ray": info["samples_per_ray"],
               "same_schedule_as_gpu_arm": info["same_schedule"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload, "name": args.config, "engine": {k: v for k, v in ekw.items()}, "n_params": int(eng.n_params),
                       "rays_per_step_per_gpu": R, "parallelism": f"ray-sharded dp{world}" + ("" if world == 1 else (" + fused NVLink exchange kernel (reduce-scatter + Adam + all-gather, "
                                                                    + ("NVSwitch multimem" if eng._symm.multicast else "peer loads/stores") + ")") if eng._symm is not None
                                                                    else " + NCCL reduce-scatter / sharded Adam / all-gather"),
                       "density_grid_update_every": 16, "optimizer": "fused Adam eps=1e-15 inside the timed region", "cuda_graph": not args.no_graph,
                       "extra_untimed_warmup_steps": extra_warmup,
                       "l2": f"per-step working set (fp32 params+grads+Adam moments {16 * eng.n_params / 1e6:.0f} MB, fp16 table {2 * eng.n_params / 1e6:.0f} MB, sample "
                             f"arrays) exceeds the 126 MB L2; {pool_n} distinct ray batches are cycled",
                       "mixed_precision": "fp16 table/weights/activations, fp32 accumulation, fp32 master params (reference: AMP precision=16)"},
            "samples_per_sec": samples_dev / (ms_dev * 1e-3), "samples_per_ray": samples_dev / rays_total,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(3 * R * 3 * 4), "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e / K,
                    "api": "NGPEngine.train_step_packed(host_pinned_batch) -> C ABI; loss copied to pinned host memory every step"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "l2_roofline": l2_roofline, "kernel_us": breakdown, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "frozen_api": frozen, "render": render,
            "final_loss_terms": final_loss, "amp": amp,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
