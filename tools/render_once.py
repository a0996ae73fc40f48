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
pose = syn.camera_poses(2, seed=7)
o, d = syn.image_rays(pose[1]); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
eng.render(o, d, min_chunk=8); torch.cuda.synchronize()
torch.cuda.profiler.start()
out = eng.render(o, d, min_chunk=8, iterations_per_batch=64)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(out["iterations"], out["total_samples"] / 640000)
