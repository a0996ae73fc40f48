import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch, time
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
for warm in (True, False):
    for _ in range(3): eng.update_density_grid(warmup=warm)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(20): eng.update_density_grid(warmup=warm)
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"update_density_grid(warmup={warm}): gpu {a.elapsed_time(b) / 20:.3f} ms, host enqueue {(t1 - t0) * 1e3 / 20:.3f} ms")
