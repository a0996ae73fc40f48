#!/usr/bin/env python
"""kernel-time breakdown of one 800x800 render (torch.profiler / CUPTI): python tools/render_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200 import engine as E
dev = torch.device("cuda", 0)
eng = E.NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, bench.R_PER_GPU, seed=1000)).to(dev)
for s in range(1, 300):
    eng.train_step_packed(pool[s % 8], global_step=s)
eng.flush()
pose = syn.camera_poses(2, seed=7)
o, d = syn.image_rays(pose[1]); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
eng.render(o, d, min_chunk=bench.RENDER_MIN_CHUNK); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = eng.render(o, d, min_chunk=bench.RENDER_MIN_CHUNK); b.record(); torch.cuda.synchronize()
print(f"frame {a.elapsed_time(b):.2f} ms, iterations {out['iterations']}, samples/ray {out['total_samples']/640000:.1f}, rows/ray {out['field_rows']/640000:.1f}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.render(o, d, min_chunk=bench.RENDER_MIN_CHUNK); torch.cuda.synchronize()
rows = sorted(((e.key[:70], e.count, e.device_time_total) for e in prof.key_averages()), key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
for k, c, t in rows[:12]:
    print(f"{k:72s} {c:5d} {t/1e3:9.3f} ms {100*t/tot:5.1f}%")
