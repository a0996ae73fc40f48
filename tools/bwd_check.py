"""How close do the field kernels' parameter gradients sit to the parity tolerance of tests/test_engine_gpu.py?  Prints, for several seeds and
shapes, max over the large entries of err / (8 % |want| + 0.2 % max|want|) and max err / max|want|, plus a bitwise determinism check of two runs.
    python tools/bwd_check.py            (fused tcgen05 kernels)
    MFN_FIELD_IMPL=v1 python tools/bwd_check.py   (unfused mma.sync pipeline, same tolerance, for comparison)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import scenes
from mfnerf_b200._lib import call, ptr, stream_ptr
from mfnerf_b200.engine import NGPEngine
from oracle import field_ref as fr


def run(seed, N, rgb_channels=64, rgb_layers=2, grid="Hash", n_tables=1):
    eng = NGPEngine(scale=0.5, n_rays=512, sample_capacity=512 * 160, log2_T=15, rgb_channels=rgb_channels, rgb_layers=rgb_layers, grid=grid, n_tables=n_tables)
    with torch.no_grad():
        eng.params[eng.n_mlp1:eng.n_xyz].uniform_(-0.5, 0.5)
        eng.params_h.copy_(eng.params)
    ref = fr.NGPRef(0.5, log2_T=15, grid=grid, n_tables=n_tables, rgb_channels=rgb_channels, rgb_layers=rgb_layers,
                    params=(eng.params[:eng.n_xyz].cpu(), eng.params[eng.off_rgb:eng.off_rgb + eng.n_rgb].cpu())).cuda()
    g = torch.Generator().manual_seed(seed)
    x = ((torch.rand(N, 3, generator=g) - 0.5)).cuda(); d = torch.randn(N, 3, generator=g).cuda()
    cfg = ctypes.byref(eng.cfg)
    cap = eng.cap
    n_dev = torch.tensor([N], dtype=torch.int32, device="cuda")
    xs = torch.zeros(cap, 3, device="cuda"); ds = torch.ones(cap, 3, device="cuda"); xs[:N] = x; ds[:N] = d
    sig_c = torch.zeros(cap, device="cuda"); rgb_c = torch.zeros(cap, 3, device="cuda")
    gs = torch.zeros(cap, device="cuda"); gc = torch.zeros(cap, 3, device="cuda")
    gs[:N] = torch.randn(N, generator=g).cuda() * 1e-2; gc[:N] = torch.randn(N, 3, generator=g).cuda() * 1e-2
    outs = []
    for rep in range(2):
        call("mfn_field_fwd", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(xs), ptr(ds), cap, ptr(n_dev), ptr(sig_c), ptr(rgb_c),
             ptr(eng.field_ws), eng.field_ws.numel(), stream_ptr())
        eng.grads.zero_(); eng.overflow.zero_()
        call("mfn_field_bwd", cfg, ptr(eng.xyz_params_h), ptr(eng.rgb_params_h), ptr(xs), cap, ptr(n_dev), ptr(gs), ptr(gc), 128.0, ptr(eng.grads),
             ptr(eng.grads[eng.off_rgb:]), ptr(eng.overflow), ptr(eng.field_ws), eng.field_ws.numel(), stream_ptr())
        torch.cuda.synchronize()
        outs.append(eng.grads.clone())
    mlp_same = torch.equal(outs[0][:eng.n_mlp1], outs[1][:eng.n_mlp1]) and torch.equal(outs[0][eng.off_rgb:], outs[1][eng.off_rgb:])
    sig_r, rgb_r = ref(x, d)
    ((sig_r * gs[:N]).sum() + (rgb_r * gc[:N]).sum()).backward()
    res = []
    for name, got, want in (("xyz", outs[0][:eng.n_xyz] / 128.0, ref.xyz_params.grad), ("rgb", outs[0][eng.off_rgb:eng.off_rgb + eng.n_rgb] / 128.0, ref.rgb_params.grad)):
        sc = want.abs().max().item()
        err = (got - want).abs()
        big = want.abs() > 5e-2 * sc
        ratio = (err[big] / (8e-2 * want.abs()[big] + 2e-3 * sc)).max().item()
        res.append(f"{name}: big-entry err/bound {ratio:.2f}, max err/scale {err.max().item() / sc:.4f}")
    return mlp_same, res


if __name__ == "__main__":
    print("impl:", os.environ.get("MFN_FIELD_IMPL", "fused"))
    for shape in ((64, 2, "Hash", 1), (128, 2, "Hash", 1), (128, 2, "MixedFeature", 8)):
        for seed in (5, 6, 7, 8):
            same, res = run(seed, 3001, *shape)
            print(shape, "seed", seed, "| MLP gradients bitwise equal over two runs:", same, "|", " ; ".join(res), flush=True)
