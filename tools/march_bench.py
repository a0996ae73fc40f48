#!/usr/bin/env python
"""times mfn_raymarching_train on the bench workload (CUDA events), for tuning: python tools/march_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine, G, MAX_SAMPLES
from mfnerf_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda", 0)
e = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
e.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); e.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(4, bench.R_PER_GPU, seed=1000)).to(dev)
e.rays.copy_(pool[1]); st = stream_ptr(dev); R = e.n_rays
call("mfn_ray_aabb_intersect", ptr(e.rays_o), ptr(e.rays_d), ptr(e.center), ptr(e.half_size), R, 1, 1, ptr(e.hit_cnt), ptr(e.hits_t), ptr(e.hits_idx), st)
call("mfn_clamp_near", ptr(e.hits_t), R, 0.01, st)
e.noise.uniform_(0, 1)
def run():
    call("mfn_raymarching_train", ptr(e.rays_o), ptr(e.rays_d), ptr(e.hits_t), ptr(e.density_bitfield), e.cascades, e.scale, e.esf, ptr(e.noise), G, MAX_SAMPLES, R,
         e.cap, ptr(e.rays_a), ptr(e.xyzs), ptr(e.dirs), ptr(e.deltas), ptr(e.ts), ptr(e.counter), ptr(e.march_ws), e.march_ws.numel(), st)
for _ in range(5): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): run()
b.record(); torch.cuda.synchronize()
import hashlib
print(hashlib.md5(e.rays_a.cpu().numpy().tobytes()).hexdigest()[:8], float(e.noise.sum()), end=" ")
print(f"K={os.environ.get('MFN_MARCH_K')} RPG={os.environ.get('MFN_MARCH_RPG', 'default')}: raymarching_train {a.elapsed_time(b) / 50 * 1e3:.1f} us, samples {int(e.counter[0])}")
