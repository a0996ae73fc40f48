"""fixed per-step cost: the bench engine with very few rays (all kernels nearly empty): what remains is launch gaps + Adam + grid update"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
for R in (64, 8192):
    eng = NGPEngine(scale=0.5, n_rays=R, device=dev, seed=1337)
    eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
    pool = torch.from_numpy(bench.make_pool(8, R, seed=1000)).to(dev)
    for s in range(1, 4): eng.train_step_packed(pool[s % 8], global_step=s)
    eng.capture()
    for s in range(4, 40): eng.train_step_packed(pool[s % 8], global_step=s)
    torch.cuda.synchronize()
    for skip_density in (False, True):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(400, 800):
            eng.train_step_packed(pool[s % 8], global_step=(s * 16 + 1) if skip_density else s)
        b.record(); torch.cuda.synchronize()
        print(f"R={R} skip_density={skip_density}: {a.elapsed_time(b) / 400 * 1e3:.1f} us/step, samples {int(eng.counter[0])}")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): eng.update_density_grid(warmup=False)
    b.record(); torch.cuda.synchronize()
    print(f"R={R}: update_density_grid after training: {a.elapsed_time(b) / 10:.3f} ms; occupied cells {int((eng.density_grid > 5.9).sum())}, positive {int((eng.density_grid > 0).sum())}")
