"""BASELINE.json config 4 shape: scale 16 (6 cascades), T = 2^21 (97 MiB fp16 table, not L2-resident), 8192 rays: sanity + step time"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn
from mfnerf_b200.engine import NGPEngine
dev = torch.device("cuda", 0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 21
eng = NGPEngine(scale=16.0, log2_T=T, n_rays=8192, device=dev, seed=1337, distortion_w=1e-3)
print("params", eng.n_params, "cascades", eng.cascades)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(16.0, 6)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, 8192, seed=1000)).to(dev)
for s in range(1, 4): eng.train_step_packed(pool[s % 8], global_step=s)
eng.capture()
for s in range(4, 40): eng.train_step_packed(pool[s % 8], global_step=s)
torch.cuda.synchronize()
l0 = eng.loss_terms.tolist()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in range(40, 240): eng.train_step_packed(pool[s % 8], global_step=s)
b.record(); torch.cuda.synchronize()
print(f"T=2^{T}: {a.elapsed_time(b) / 200:.3f} ms/step, {8192 * 200 / a.elapsed_time(b) * 1e3 / 1e6:.2f} M rays/s, samples/step {int(eng.counter[0])}, loss {l0} -> {eng.loss_terms.tolist()}, overflow {int(eng.overflow[0])}")
