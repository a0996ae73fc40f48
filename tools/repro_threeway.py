"""Diagnosis of tests/test_engine_gpu.py::test_pipelined_and_data_parallel_step_paths_reproduce_the_plain_path (red on the driver's box in
round 1): runs the three step paths N times each and prints, per pair, how many parameters differ by more than the test's bound, where
they live (sigma MLP | hash table by level | rgb MLP) and how two runs of the SAME path differ (the float-atomic noise floor).
    python tools/repro_threeway.py [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import scenes
from mfnerf_b200 import field_ops


def engine(n_rays, **kw):
    import vren
    from mfnerf_b200.engine import NGPEngine
    eng = NGPEngine(scale=0.5, n_rays=n_rays, sample_capacity=n_rays * 160, log2_T=15, **kw)
    eng.density_grid.copy_(torch.from_numpy(scenes.syn.lego_density_grid(0.5, 1)).cuda())
    vren.packbits(eng.density_grid.reshape(-1), 0.5, eng.density_bitfield)
    return eng


def run(kw, batches, noise, steps=12):
    eng = engine(256, **kw)
    eng.fixed_noise = noise
    losses = []
    for s in range(1, 4):
        eng.train_step_packed(batches[s % 3], global_step=s)
    eng.capture()
    for s in range(4, steps + 1):
        eng.train_step_packed(batches[s % 3], global_step=s)
        eng.flush(); losses.append(eng.loss_terms.clone())
    p = eng.gather_master_params().clone()
    torch.cuda.synchronize()
    return eng, p, torch.stack(losses).cpu(), int(eng.counter[0])


def describe(eng, a, b, tag):
    d = (a - b).abs()
    scale = a.abs().max().item()
    bad = d > 2e-3 * scale
    entries, offs, *_ = field_ops.grid_layout(eng.cfg.grid)
    regions = [("sigma_mlp", 0, eng.n_mlp1)]
    try:
        offs = list(offs)
        for l in range(len(offs) - 1):
            regions.append((f"L{l}", eng.n_mlp1 + 2 * int(offs[l]), eng.n_mlp1 + 2 * int(offs[l + 1])))
    except Exception:
        regions.append(("grid", eng.n_mlp1, eng.n_xyz))
    regions.append(("rgb_mlp", eng.off_rgb, eng.off_rgb + eng.n_rgb))
    where = {n: int(bad[lo:hi].sum()) for n, lo, hi in regions if int(bad[lo:hi].sum())}
    print(f"  {tag}: max|d|={d.max().item():.3e} (bound {2e-3 * scale:.3e}, scale {scale:.3f}) n_bad={int(bad.sum())} of {a.numel()}  where={where}", flush=True)
    return int(bad.sum())


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29544", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    batches = []
    for k in range(3):
        rays = scenes.scene("lego", 256, seed=11 + k)
        tgt = scenes.syn.analytic_render(rays["rays_o"], rays["rays_d"]).float()
        batches.append(torch.stack([torch.from_numpy(rays["rays_o"]), torch.from_numpy(rays["rays_d"]), tgt]).cuda().contiguous())
    noise = torch.rand(256, device="cuda", generator=torch.Generator("cuda").manual_seed(6))
    variants = [("plain", dict(pipelined=False)), ("plain2", dict(pipelined=False)), ("pipelined", dict(pipelined=True)), ("dp", dict(force_dp_path=True))]
    tot = {}
    for r in range(reps):
        print(f"repeat {r}", flush=True)
        res = {}
        for name, kw in variants:
            res[name] = run(kw, batches, noise)
        eng0, p0, l0, c0 = res["plain"]
        for name in ("plain2", "pipelined", "dp"):
            eng, p, l, c = res[name]
            nb = describe(eng0, p0, p, f"plain vs {name}")
            tot[name] = tot.get(name, 0) + (1 if nb else 0)
            print(f"     samples {c0} vs {c}; max rel loss diff over steps {((l0 - l).abs() / l0.abs().clamp_min(1e-6)).max().item():.3e}", flush=True)
    print("runs with any out-of-bound entry:", tot, "of", reps)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
