#!/usr/bin/env python
"""Experiment: how well do the hash-grid scatter (L1/L2 reduction bound) and Adam (HBM bound) overlap when they run at the same time?
The optimiser runs on a SECOND parameter set of the same size, so the two have no dependency.  Times: backward+scatter alone, Adam alone,
both back to back on one stream, both on two streams."""
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
print("samples", n)
cfg = ctypes.byref(eng.cfg)
eng.n_field.fill_(n)
P = eng.n_params
p2, g2, m2, v2 = (torch.zeros(P, device=dev) for _ in range(4))
h2 = torch.zeros(P, device=dev, dtype=torch.float16)
g2.normal_()
s0 = torch.cuda.current_stream(dev)
sA, sB = torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev)
S = eng.cap
call("mfn_field_fwd", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(eng.xyzs), ptr(eng.dirs), S, ptr(eng.n_field), ptr(eng.sigmas), ptr(eng.rgbs),
     ptr(eng.field_ws), eng.field_ws.numel(), ctypes.c_void_p(s0.cuda_stream))


def bwd(st):
    call("mfn_field_bwd_amp", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(eng.xyzs), S, ptr(eng.n_field), ptr(eng.dL_dsigmas), ptr(eng.dL_drgbs),
         ptr(eng._amp), ptr(eng.grads), ptr(eng.grads[eng.off_rgb:]), ptr(eng.overflow), ptr(eng.field_ws), eng.field_ws.numel(), ctypes.c_void_p(st.cuda_stream))


def adam(st):
    call("mfn_adam_step_amp", ptr(p2), ptr(g2), ptr(m2), ptr(v2), ptr(h2), P, ptr(eng._adam_hyper), 0.9, 0.999, 1e-15, 1.0, ptr(eng._amp), ptr(eng.overflow), 0,
         ctypes.c_void_p(st.cuda_stream))


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


def both_seq():
    bwd(s0); adam(s0)


def both_par():
    sA.wait_stream(s0); sB.wait_stream(s0)
    bwd(sA); adam(sB)
    s0.wait_stream(sA); s0.wait_stream(sB)


mid = torch.cuda.Event()


def par_after_bwd():
    # Adam starts when the MLP backward kernel has finished, i.e. overlaps the scatter only: approximated by delaying it with a sleep-free
    # dependency on an event recorded after a first backward-only launch is not possible through the C ABI (one call = backward + scatter),
    # so this variant simply launches Adam second on the low-priority stream
    sA.wait_stream(s0); sB.wait_stream(s0)
    bwd(sA); adam(sB)
    s0.wait_stream(sA); s0.wait_stream(sB)


for name, fn in (("backward + scatter", lambda: bwd(s0)), ("adam", lambda: adam(s0)), ("back to back", both_seq), ("two streams", both_par), ("back to back", both_seq)):
    eng.grads.zero_()
    print(f"{name:24s} {timed(fn):8.1f} us")
