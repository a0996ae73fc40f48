#!/usr/bin/env python
"""phase timeline of CTA 0 of the fused forward kernel (clock64 stamps written through FusedArgs.dbg)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
dbg = torch.zeros(256 + 6 * 40, dtype=torch.int64, device="cuda")
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

# backward kernel: stamps after every MMA-completion wait ("w") and every barrier ("s") of CTA 0 / thread 0
b = dbg.cpu()[256:].view(6, 40).numpy()
print("backward kernel, cycles between consecutive stamps (tile start, load+sync, then wait/sync per stage):")
for k in range(6):
    r = [int(v) for v in b[k] if v != 0]
    if len(r) < 3: continue
    print(f"tile {k}: total {r[-1]-r[0]:7d} | " + " ".join(str(r[j] - r[j - 1]) for j in range(1, len(r))))
