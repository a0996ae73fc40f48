#!/usr/bin/env python
"""Experiment: does the hash-grid scatter of one half of the samples overlap the MLP backward of the other half?
(a) one call over all samples; (b) two half-size calls back to back on one stream; (c) the two halves on two streams."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from mfnerf_b200 import synthetic as syn, _lib
from mfnerf_b200.engine import NGPEngine, ptr, call

dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, bench.R_PER_GPU, seed=1000)).to(dev)
for s in range(1, 40):
    eng.train_step_packed(pool[s % 8], global_step=s)
eng.flush(); torch.cuda.synchronize()
n = int(eng.counter[0].item())
h = (n // 2) // 128 * 128
print("samples", n, "half", h)
cfg = ctypes.byref(eng.cfg)
xyz, dirs, ds, dr = eng.xyzs[:n].clone(), eng.dirs[:n].clone(), eng.dL_dsigmas[:n].clone(), eng.dL_drgbs[:n].clone()
sig, rgb = torch.empty(n, device=dev), torch.empty(n, 3, device=dev)


class Part:
    def __init__(self, lo, hi):
        self.n = hi - lo
        self.xyz, self.dirs, self.ds, self.dr = xyz[lo:hi].contiguous(), dirs[lo:hi].contiguous(), ds[lo:hi].contiguous(), dr[lo:hi].contiguous()
        self.sig, self.rgb = torch.empty(self.n, device=dev), torch.empty(self.n, 3, device=dev)
        self.ws = torch.empty(_lib.lib.mfn_field_workspace_bytes(cfg, self.n, 1), dtype=torch.uint8, device=dev)
        self.cnt = torch.full((1,), self.n, dtype=torch.int32, device=dev)

    def fwd(self, st):
        call("mfn_field_fwd", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(self.xyz), ptr(self.dirs), self.n, ptr(self.cnt), ptr(self.sig), ptr(self.rgb),
             ptr(self.ws), self.ws.numel(), st)

    def bwd(self, st):
        call("mfn_field_bwd_amp", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(self.xyz), self.n, ptr(self.cnt), ptr(self.ds), ptr(self.dr),
             ptr(eng._amp), ptr(eng.grads), ptr(eng.grads[eng.off_rgb:]), ptr(eng.overflow), ptr(self.ws), self.ws.numel(), st)


full, a, b = Part(0, n), Part(0, h), Part(h, n)
s0 = torch.cuda.current_stream(dev)
sA, sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
for p in (full, a, b):
    p.fwd(ctypes.c_void_p(s0.cuda_stream))
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=8):
    ts = []
    for _ in range(reps):
        e0, e1 = E(), E()
        torch.cuda.synchronize()
        e0.record(s0); fn(); e1.record(s0)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[2:]))


def one():
    full.bwd(ctypes.c_void_p(s0.cuda_stream))


def seq():
    a.bwd(ctypes.c_void_p(s0.cuda_stream)); b.bwd(ctypes.c_void_p(s0.cuda_stream))


def par():
    sA.wait_stream(s0); sB.wait_stream(s0)
    a.bwd(ctypes.c_void_p(sA.cuda_stream)); b.bwd(ctypes.c_void_p(sB.cuda_stream))
    s0.wait_stream(sA); s0.wait_stream(sB)


for name, fn in (("one call", one), ("two halves, one stream", seq), ("two halves, two streams", par), ("one call", one)):
    eng.grads.zero_()
    print(f"{name:28s} {timed(fn):8.1f} us")
