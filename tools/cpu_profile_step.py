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
pr = cProfile.Profile(); pr.enable()
for s in range(300, 620):
    eng.train_step_packed(pool[s % 8], global_step=s)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
