"""Timing of the device-side ray generation / batch sampling (mfn_ray_batch), mark_invisible_cells, and of a training step that draws
its batch from a data set resident in HBM versus the same step fed from pinned host memory (the bench's e2e arm).
    python tools/dataset_bench.py            (on the GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from mfnerf_b200 import synthetic as syn, dataset as mds
from mfnerf_b200.engine import NGPEngine

dev = torch.device("cuda", 0)
W, H = syn.IMG_WH
K = np.array([[syn.FOCAL, 0, W / 2], [0, syn.FOCAL, H / 2], [0, 0, 1]], np.float32)
N_IMG = 100
poses = syn.camera_poses(N_IMG, seed=0)


def timed(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3      # us


# the data set: 100 views of the analytic scene (768 MB fp32 in HBM)
pix = torch.empty(N_IMG, W * H, 3, device=dev)
dirs = torch.from_numpy(syn.pixel_directions(np.arange(W * H))).to(dev)
for i, p in enumerate(poses):
    P = torch.from_numpy(p).to(dev)
    pix[i] = syn.analytic_render(P[:, 3].expand(W * H, 3).contiguous(), dirs @ P[:, :3].T)
ds = mds.ResidentDataset(K, (W, H), poses, pix, device=dev)
R = bench.R_PER_GPU
ctr = torch.zeros(1, dtype=torch.int64, device=dev)
o, d, t = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev)
us = timed(lambda: mds.ray_batch(ds.camera, ds.poses, R, pixels=ds.pixels, strategy="all_images", seed=1, call_counter=ctr, rays_o=o, rays_d=d, rgb=t))
print(f"ray_batch draw+generate+target gather, {R} rays: {us:.1f} us  ({R * (48 + 12 + 36) / us / 1e3:.1f} GB/s of 96 B/ray)")
n = W * H
vo, vd, vt = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
us = timed(lambda: mds.ray_batch(ds.camera, ds.poses, n, pixels=ds.pixels, image=7, rays_o=vo, rays_d=vd, rgb=vt))
print(f"ray_batch whole 800x800 view: {us:.1f} us  ({n * 48 / us / 1e3:.0f} GB/s of 48 B/ray: 12 B pixel in, 36 B out)")

eng = NGPEngine(scale=0.5, n_rays=R, device=dev, seed=1337)
us = timed(lambda: eng.mark_invisible_cells(ds.K, ds.poses, ds.img_wh), n=10, warm=2)
print(f"mark_invisible_cells 128^3 x {N_IMG} cameras (incl. the host-side tensor prep): {us:.0f} us;  valid cells {float((eng.density_grid == 0).float().mean()):.3f}")

# training: resident data set vs pinned host batches
eng.density_grid.copy_(torch.from_numpy(syn.lego_density_grid(0.5, 1)).to(dev)); eng.repack_bitfield(0.5)
eng.attach_dataset(ds, seed=3)
for s in range(1, 4): eng.train_step_resident(global_step=s)
eng.capture()
for s in range(4, 600): eng.train_step_resident(global_step=s)
snap = eng.snapshot()
host = []
for k in range(16):
    oo, dd, tt, _, _ = mds.ray_batch(ds.camera, ds.poses, R, pixels=ds.pixels, strategy="all_images", seed=100 + k, return_indices=True)
    host.append(torch.stack([oo, dd, tt]).cpu().pin_memory())
loss = torch.zeros(3).pin_memory()
for name, step in (("resident data set (no H2D)", lambda s: eng.train_step_resident(global_step=s)),
                   ("pinned host batches (bench e2e arm)", lambda s: eng.train_step_packed(host[s % 16], global_step=s))):
    eng.restore(snap)
    for s in range(600, 700): step(s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.reset_samples_marched()
    a.record()
    for s in range(700, 1700):
        step(s); eng.loss_to_host(loss)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 1000
    print(f"train step, {name}: {ms:.4f} ms/step = {R / ms / 1e3:.2f} M rays/s, {int(eng.samples_marched()) / 1000:.0f} samples/step, loss {float(eng.loss_terms.sum()):.5f}")
