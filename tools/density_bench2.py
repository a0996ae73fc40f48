import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import torch
import bench
from mfnerf_b200 import synthetic as syn, _lib
from mfnerf_b200.engine import NGPEngine, G
from mfnerf_b200._lib import call, ptr, stream_ptr
dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
eng.update_density_grid(warmup=False); eng.update_density_grid(warmup=True)
st = stream_ptr(dev); G3 = G ** 3; M = G3 // 4
def t(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:40s} {a.elapsed_time(b) / reps * 1e3:8.1f} us")
grid_c = eng.density_grid[0]
t("positions random", lambda: call("mfn_grid_cell_positions", None, 1, M, 0, 0.5, G, 123, ptr(eng._dg_idx), ptr(eng._dg_xyz), st))
t("cumsum", lambda: torch.cumsum(grid_c > 5.9, 0, dtype=torch.int32))
cs = torch.cumsum(grid_c > 5.9, 0, dtype=torch.int32)
t("rand*cs", lambda: (torch.rand(M, device=dev) * cs[-1]).to(torch.int32))
k = (torch.rand(M, device=dev) * cs[-1]).to(torch.int32)
t("searchsorted", lambda: torch.searchsorted(cs, k, right=True).clamp_(max=G3 - 1))
def assign(): eng._dg_idx[M:] = torch.searchsorted(cs, k, right=True).clamp_(max=G3 - 1)
t("searchsorted+assign", assign)
t("positions idx", lambda: call("mfn_grid_cell_positions", ptr(eng._dg_idx[M:]), 0, M, 0, 0.5, G, 5, None, ptr(eng._dg_xyz[M:]), st))
ws = eng.field_ws
eng._dg_idx.copy_(torch.sort(eng._dg_idx)[0])
t("sort 1M", lambda: torch.sort(eng._dg_idx)[0])
call("mfn_grid_cell_positions", ptr(eng._dg_idx), 0, 2 * M, 0, 0.5, G, 77, None, ptr(eng._dg_xyz), st)
print("unique cells", int(torch.unique(eng._dg_idx).numel()))
t("density 1M sampled", lambda: call("mfn_density_fwd", ctypes.byref(eng.cfg), ptr(eng.xyz_params_h), ptr(eng._dg_xyz), 2 * M, None, ptr(eng._dg_sig), ptr(ws), ws.numel(), st))
t("grid_update sampled", lambda: call("mfn_grid_update", ptr(grid_c), ptr(eng._dg_idx), ptr(eng._dg_sig), G3, 2 * M, 0.95, st))
t("mean", lambda: call("mfn_grid_mean_positive", ptr(eng.density_grid), G3, ptr(eng._dg_scratch), ptr(eng._dg_mean), st))
call("mfn_grid_cell_positions", None, 0, G3, 0, 0.5, G, 9, None, ptr(eng._dg_xyz), st)
t("density 2M all cells", lambda: call("mfn_density_fwd", ctypes.byref(eng.cfg), ptr(eng.xyz_params_h), ptr(eng._dg_xyz), G3, None, ptr(eng._dg_sig), ptr(ws), ws.numel(), st))
t("grid_update all", lambda: call("mfn_grid_update", ptr(grid_c), None, ptr(eng._dg_sig), G3, G3, 0.95, st))
