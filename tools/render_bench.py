#!/usr/bin/env python
"""800x800 render timing with different minimum chunk sizes: python tools/render_bench.py"""
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for s in range(1, steps):
    eng.train_step_packed(pool[s % 8], global_step=s)
pose = syn.camera_poses(2, seed=7)
o, d = syn.image_rays(pose[1]); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
for esf_min in (1, 2, 4, 8):
    for ipb in (8, 32):
        eng.esf = 0.0
        # monkeypatch the minimum chunk: engine derives it from esf; use the private knob
        E_min = esf_min
        def run():
            old = eng.esf
            out = eng.render(o, d, iterations_per_batch=ipb) if E_min == 1 else eng.render(o, d, iterations_per_batch=ipb, min_chunk=E_min)
            return out
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = run(); b.record(); torch.cuda.synchronize()
        print(f"min_chunk={esf_min} ipb={ipb}: {a.elapsed_time(b):.2f} ms, iterations {out['iterations']}, samples/ray {out['total_samples']/640000:.1f}, rows/ray {out['field_rows']/640000:.1f}")
