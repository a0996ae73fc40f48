import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, bench.R_PER_GPU, seed=1000)).to(dev)
for s in range(1, 4):
    eng.train_step_packed(pool[s % 8], global_step=s)
eng.capture()
for s in range(4, 300):
    eng.train_step_packed(pool[s % 8], global_step=s)
torch.cuda.synchronize()
pool_h = pool.cpu().pin_memory()
loss_h = torch.zeros(400, 3).pin_memory()
for s in range(300, 340):
    eng.train_step_packed(pool_h[s % 8], global_step=s); eng.loss_to_host(loss_h[s - 300])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for s in range(340, 660):
    eng.train_step_packed(pool_h[s % 8], global_step=s); eng.loss_to_host(loss_h[s - 340])
pr.disable(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host loop {(t1 - t0) / 320 * 1e3:.3f} ms/step, incl. drain {(t2 - t0) / 320 * 1e3:.3f} ms/step")
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
