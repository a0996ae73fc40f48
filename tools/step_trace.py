import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
R = 8192
eng = NGPEngine(scale=0.5, n_rays=R, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, R, seed=1000)).to(dev)
for s in range(1, 4): eng.train_step_packed(pool[s % 8], global_step=s)
eng.capture()
for s in range(4, 400): eng.train_step_packed(pool[s % 8], global_step=s)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
host = []
evs[0].record()
for k, s in enumerate(range(401, 449)):
    t0 = time.perf_counter()
    eng.train_step_packed(pool[s % 8], global_step=s)
    host.append((time.perf_counter() - t0) * 1e3)
    evs[k + 1].record()
torch.cuda.synchronize()
for k, s in enumerate(range(401, 449)):
    print(f"step {s} {'UPDATE' if s % 16 == 0 else '      '} gpu {evs[k].elapsed_time(evs[k + 1]):.3f} ms  host {host[k]:.3f} ms")
