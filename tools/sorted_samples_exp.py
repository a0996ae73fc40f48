#!/usr/bin/env python
"""Experiment: how much faster are the table kernels (field_fwd gather, grid scatter) when the samples of a step are presented in
spatial (Morton) order instead of ray order?  Same samples, same kernels, only the row order differs.
usage: python tools/sorted_samples_exp.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from mfnerf_b200 import synthetic as syn, _lib
from mfnerf_b200.engine import NGPEngine, ptr, stream_ptr, call

dev = torch.device("cuda", 0)
eng = NGPEngine(scale=0.5, n_rays=bench.R_PER_GPU, device=dev, seed=1337)
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
pool = torch.from_numpy(bench.make_pool(8, bench.R_PER_GPU, seed=1000)).to(dev)
for s in range(1, 40):
    eng.train_step_packed(pool[s % 8], global_step=s)
eng.flush(); torch.cuda.synchronize()
n = int(eng.counter[0].item())
print("samples", n)
xyz0, dir0 = eng.xyzs[:n].clone(), eng.dirs[:n].clone()
ds0, dr0 = eng.dL_dsigmas[:n].clone(), eng.dL_drgbs[:n].clone()


def part1by2(v):
    v = v & 0x3ff
    v = (v | (v << 16)) & 0x30000ff
    v = (v | (v << 8)) & 0x300f00f
    v = (v | (v << 4)) & 0x30c30c3
    v = (v | (v << 2)) & 0x9249249
    return v


def morton(xyz, res):
    q = ((xyz / (2 * eng.scale) + 0.5).clamp(0, 1 - 1e-6) * res).long()
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


def order(kind):
    if kind == "ray":
        return torch.arange(n, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    shuffle = torch.randperm(n, device=dev, generator=g)
    res = {"morton1024": 1024, "bin128": 128, "bin64": 64, "bin32": 32, "bin256": 256}[kind]
    key = morton(xyz0[shuffle], res)               # random order inside a bin (what an atomic counting sort gives)
    return shuffle[torch.sort(key, stable=True)[1]]


cfg = ctypes.byref(eng.cfg)
S = eng.cap
st = stream_ptr(dev)
for kind in ("ray", "morton1024", "bin256", "bin128", "bin64", "bin32", "ray"):
    perm = order(kind)
    eng.xyzs[:n].copy_(xyz0[perm]); eng.dirs[:n].copy_(dir0[perm])
    eng.dL_dsigmas[:n].copy_(ds0[perm]); eng.dL_drgbs[:n].copy_(dr0[perm])
    eng.n_field.fill_(n)
    names = ("field_fwd", "field_bwd", "grid_encode_bwd")
    evs = {}
    for name in names:
        a, b = _lib.lib.mfn_event_create(), _lib.lib.mfn_event_create()
        evs[name] = (a, b)
        _lib.check(_lib.lib.mfn_profile_set(name.encode(), a, b), "mfn_profile_set")
    acc = {k: [] for k in names}
    for it in range(6):
        call("mfn_field_fwd", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(eng.xyzs), ptr(eng.dirs), S, ptr(eng.n_field), ptr(eng.sigmas),
             ptr(eng.rgbs), ptr(eng.field_ws), eng.field_ws.numel(), st)
        eng.overflow.zero_()
        call("mfn_field_bwd_amp", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(eng.xyzs), S, ptr(eng.n_field), ptr(eng.dL_dsigmas), ptr(eng.dL_drgbs),
             ptr(eng._amp), ptr(eng.grads), ptr(eng.grads[eng.off_rgb:]), ptr(eng.overflow), ptr(eng.field_ws), eng.field_ws.numel(), st)
        torch.cuda.synchronize()
        ms = ctypes.c_float()
        for name, (a, b) in evs.items():
            if _lib.lib.mfn_event_elapsed_ms(a, b, ctypes.byref(ms)) == 0:
                acc[name].append(ms.value * 1e3)
    _lib.lib.mfn_profile_set(b"", None, None)
    print(f"{kind:11s}", {k: round(float(np.median(v[1:])), 1) for k, v in acc.items() if v}, "us; grads sum", float(eng.grads.double().abs().sum()))
    eng.grads.zero_()
