#!/usr/bin/env python
"""phase timeline of CTA 0 of the fused forward kernel (clock64 stamps written through FusedArgs.dbg)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
dbg = torch.zeros(1024 + 4 * 4096, dtype=torch.int64, device="cuda")
os.environ["MFN_FWD_DBG"] = str(dbg.data_ptr())
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, bench.R_PER_GPU, seed=1000)).to(dev)
for s in range(1, 20):
    eng.train_step_packed(pool[s % 8], global_step=s)
torch.cuda.synchronize()
t = dbg.cpu()[:192].view(12, 16).numpy()
names = ["start", "gather_done(t0)", "sync1", "mma1_done", "epi1", "sync2", "epi2(sigma)", "sync3", "tile_end"]
print("cycles per phase (CTA 0, thread 0; gather_done of thread 255 in last column)")
for k in range(12):
    r = t[k]
    if r[8] == 0: continue
    d = [int(r[j] - r[j - 1]) for j in range(1, 9)]
    print(f"tile {k:2d}: total {int(r[8]-r[0]):7d} | " + " ".join(f"{names[j]}={d[j-1]}" for j in range(1, 9)) + f" | t255 gather={int(r[9]-r[0])}")

import numpy as np
e = dbg.cpu()[1024:].view(-1, 4).numpy()
e = e[e[:, 1] > 0]
t0 = e[:, 0].min()
dur = (e[:, 1] - e[:, 0]) / 1e3
print(f"forward kernel, {len(e)} CTAs: entry {((e[:,0]-t0)/1e3).min():.1f}..{((e[:,0]-t0)/1e3).max():.1f} us, exit {((e[:,1]-t0)/1e3).min():.1f}..{((e[:,1]-t0)/1e3).max():.1f} us, "
      f"CTA duration min {dur.min():.1f} mean {dur.mean():.1f} max {dur.max():.1f} us")
per_tile = dur / np.maximum(e[:, 3], 1)
sm = e[:, 2]
by_sm = np.array([per_tile[sm == k].mean() for k in np.unique(sm)])
print("  us per tile by SM (sorted):", " ".join(f"{v:.1f}" for v in np.sort(by_sm)[::8]))
print("  exit time by SM (sorted):", " ".join(f"{v:.0f}" for v in np.sort(np.array([((e[:,1]-t0)/1e3)[sm == k].max() for k in np.unique(sm)]))[::8]))
print("samples of the last step", int(eng.counter[0].item()))

# backward kernel: stamps after every MMA-completion wait ("w") and every barrier ("s") of CTA 0 / thread 0
b = dbg.cpu()[256:496].view(6, 40).numpy()
print("backward kernel, cycles between consecutive stamps (tile start, load+sync, then wait/sync per stage):")
for k in range(6):
    r = [int(v) for v in b[k] if v != 0]
    if len(r) < 3: continue
    print(f"tile {k}: total {r[-1]-r[0]:7d} | " + " ".join(str(r[j] - r[j - 1]) for j in range(1, len(r))))
