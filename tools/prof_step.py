#!/usr/bin/env python
"""One eager training step of the bench workload bracketed by cudaProfilerStart/Stop, for ncu:
   ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:<kernels> -o gpurun_out/x python tools/prof_step.py
Also usable without ncu as a quick sanity run (prints the loss terms)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from mfnerf_b200 import synthetic as syn  # noqa: E402
from mfnerf_b200.engine import NGPEngine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev))
eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, bench.R_PER_GPU, seed=1000)).to(dev)
for s in range(1, steps):
    eng.train_step_packed(pool[s % 8], global_step=s)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step_packed(pool[steps % 8], global_step=steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("samples", int(eng.counter[0].item()), "loss", eng.loss_terms.tolist())
