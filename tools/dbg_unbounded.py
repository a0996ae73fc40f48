import os, sys
os.environ["MFN_DEBUG_SYNC"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import scenes
from mfnerf_b200.engine import NGPEngine
eng = NGPEngine(scale=16.0, n_rays=1024, sample_capacity=1024 * 256, log2_T=17, distortion_w=1e-3)
eng.density_grid.copy_(torch.from_numpy(scenes.syn.lego_density_grid(16.0, 6)).cuda()); eng.repack_bitfield(0.5)
o, d, _, _ = scenes.syn.random_rays(1024, seed=21)
tgt = scenes.syn.analytic_render(o, d)
o = torch.from_numpy(o).cuda(); d = torch.from_numpy(d).cuda(); tgt = tgt.cuda().float()
def sync(tag):
    torch.cuda.synchronize(); print("ok", tag, flush=True)
for s in range(1, 40):
    if s % 16 == 0:
        eng.update_density_grid(warmup=True); sync(f"{s} density")
    eng.rays_o.copy_(o); eng.rays_d.copy_(d); eng.target.copy_(tgt)
    eng._march(); sync(f"{s} march n={int(eng.counter[0])}")
    eng._field_backward(); sync(f"{s} fb")
    eng._optimizer_step(); sync(f"{s} adam")
