"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference hot path executed on the host cores.

vren and tiny-cuda-nn have no CPU kernels, so "the reference on CPU" is a port: the plain-C restatement of the vren kernels
(oracle/vren_oracle.c, pinned bit-exact against the reference's own kernels) + the torch restatement of tcnn's
encoder / SH / fused MLPs (oracle/field_ref.py, parity unpinned), glued exactly like the reference's python layer:
  train.py:164-190 training_step  ->  rendering.py:121-163 __render_rays_train  ->  losses.py:47-60 NeRFLoss  ->  backward  ->  Adam(eps=1e-15).
Used by bench.py (`cpu_baseline`, `--impl reference`) and by tests; never by the product.
"""
import numpy as np
import torch

from . import field_ref as fr
from . import vren_oracle as orc

NEAR = 0.01
MAX_SAMPLES = 1024
G = 128


class CpuTrainer:
    def __init__(self, scale=0.5, log2_T=19, rgb_channels=64, rgb_layers=2, lr=1e-2, seed=1337, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        self.scale = scale
        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)
        self.esf = 1.0 / 256 if scale > 0.5 else 0.0
        self.model = fr.NGPRef(scale, log2_T=log2_T, rgb_channels=rgb_channels, rgb_layers=rgb_layers)
        with torch.no_grad():   # tcnn-style init: Xavier-uniform MLPs, U(-1e-4, 1e-4) grid
            g = torch.Generator().manual_seed(seed)
            n_mlp = 64 * 32 + 16 * 64
            self.model.xyz_params[:n_mlp].uniform_(-0.25, 0.25, generator=g)
            self.model.xyz_params[n_mlp:].uniform_(-1e-4, 1e-4, generator=g)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, eps=1e-15)
        self.center = np.zeros((1, 3), np.float32)
        self.half = np.full((1, 3), scale, np.float32)
        self.bitfield = np.zeros(self.cascades * G ** 3 // 8, np.uint8)
        self.rng = np.random.RandomState(seed)

    def set_density_grid(self, grid, thr=0.5):
        self.density_grid = np.ascontiguousarray(grid, np.float32).reshape(self.cascades, G ** 3).copy()
        orc.packbits(self.density_grid.reshape(-1), thr, self.bitfield)

    @torch.no_grad()
    def update_density_grid(self, density_threshold=0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=False, decay=0.95, chunk=1 << 19):
        """networks.py:242-271 (+ get_all_cells :157-168, sample_uniform_and_occupied_cells :170-197) on the host"""
        G3 = G ** 3
        tmp = np.zeros_like(self.density_grid)
        for c in range(self.cascades):
            if warmup:
                indices = np.arange(G3, dtype=np.int64)
                coords = orc.morton3D_invert(indices.astype(np.int32))
            else:
                M = G3 // 4
                coords1 = self.rng.randint(0, G, size=(M, 3)).astype(np.int32)
                indices1 = orc.morton3D(coords1).astype(np.int64)
                occ = np.nonzero(self.density_grid[c] > density_threshold)[0]
                indices2 = occ[self.rng.randint(0, len(occ), size=M)] if len(occ) > 0 else occ
                coords2 = orc.morton3D_invert(indices2.astype(np.int32))
                indices, coords = np.concatenate([indices1, indices2]), np.concatenate([coords1, coords2])
            sc = min(2.0 ** (c - 1), self.scale)
            hgs = sc / G
            xyz = (coords.astype(np.float32) / (G - 1) * 2 - 1) * (sc - hgs)
            xyz += (self.rng.rand(*xyz.shape).astype(np.float32) * 2 - 1) * hgs
            sig = np.empty(xyz.shape[0], np.float32)
            for i in range(0, xyz.shape[0], chunk):
                sig[i:i + chunk] = self.model.density_h(torch.from_numpy(xyz[i:i + chunk]))[0].numpy()
            tmp[c, indices] = sig
        g = self.density_grid
        self.density_grid = np.where(g < 0, g, np.maximum(g * decay, tmp)).astype(np.float32)
        pos = self.density_grid[self.density_grid > 0]
        mean = float(pos.mean()) if pos.size else 0.0
        orc.packbits(self.density_grid.reshape(-1), min(mean, density_threshold), self.bitfield)

    def training_step(self, global_step, rays_o, rays_d, target):
        """train.py:164-190: occupancy update every 16 steps (all cells during the first 256), then the optimisation step"""
        if global_step % 16 == 0:
            self.update_density_grid(warmup=global_step < 256)
        return self.train_step(rays_o, rays_d, target)

    def train_step(self, rays_o, rays_d, target, lambda_opacity=1e-3):
        """one optimisation step on (R,3) float32 numpy rays -> (loss, n_samples)"""
        R = rays_o.shape[0]
        _, ht, _ = orc.ray_aabb_intersect(rays_o, rays_d, self.center, self.half, 1)
        h = np.ascontiguousarray(ht[:, 0]).copy()
        m = (h[:, 0] >= 0) & (h[:, 0] < NEAR); h[m, 0] = NEAR                                   # rendering.py:29
        noise = self.rng.rand(R).astype(np.float32)
        ra, xyzs, dirs, deltas, ts, counter = orc.raymarching_train(rays_o, rays_d, h, self.bitfield, self.cascades, self.scale, self.esf, noise,
                                                                    G, MAX_SAMPLES)
        n = int(counter[0])
        self.opt.zero_grad(set_to_none=True)
        sig, rgb = self.model(torch.from_numpy(xyzs), torch.from_numpy(dirs))
        sig_np = sig.detach().numpy().astype(np.float32); rgb_np = np.ascontiguousarray(rgb.detach().numpy().astype(np.float32))
        total, opacity, depth, rgb_ray, ws = orc.composite_train_fw(sig_np, rgb_np, deltas, ts, ra, 1e-4)
        bg = 1.0 if self.esf == 0 else 0.0                                                     # rendering.py:153-161
        final = rgb_ray + bg * (1 - opacity)[:, None]
        err = final - target
        oe = opacity + 1e-10
        loss = float((err ** 2).mean() + (lambda_opacity * -oe * np.log(oe)).mean())            # losses.py:47-53, train.py:178
        g_rgb = (2 * err / err.size).astype(np.float32)
        g_op = (-(g_rgb * bg).sum(1) + lambda_opacity * -(np.log(oe) + 1) / R).astype(np.float32)
        dsig, drgbs = orc.composite_train_bw(g_op, np.zeros(R, np.float32), g_rgb, np.zeros(n, np.float32), sig_np, rgb_np, ws, deltas, ts, ra,
                                             opacity, depth, rgb_ray, 1e-4)
        torch.autograd.backward([sig, rgb], [torch.from_numpy(dsig), torch.from_numpy(drgbs)])
        self.opt.step()
        return loss, n

    @torch.no_grad()
    def render(self, rays_o, rays_d, max_samples=MAX_SAMPLES, T_thr=1e-4):
        """rendering.py:46-118 __render_rays_test -> dict(rgb, depth, opacity, total_samples)"""
        N = rays_o.shape[0]
        _, ht, _ = orc.ray_aabb_intersect(rays_o, rays_d, self.center, self.half, 1)
        h = np.ascontiguousarray(ht[:, 0]).copy()
        m = (h[:, 0] >= 0) & (h[:, 0] < NEAR); h[m, 0] = NEAR
        opacity = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); rgb = np.zeros((N, 3), np.float32)
        alive = np.arange(N, dtype=np.int64)
        samples = total = 0
        min_samples = 1 if self.esf == 0 else 4
        while samples < max_samples:
            if alive.shape[0] == 0:
                break
            ns = max(min(N // alive.shape[0], 64), min_samples)
            samples += ns
            xyzs, dirs, deltas, ts, n_eff = orc.raymarching_test(rays_o, rays_d, h, alive, self.bitfield, self.cascades, self.scale, self.esf, G,
                                                                 MAX_SAMPLES, ns)
            total += int(n_eff.sum())
            valid = ~np.all(dirs.reshape(-1, 3) == 0, axis=1)
            if not valid.any():
                break
            sig = np.zeros(valid.shape[0], np.float32); col = np.zeros((valid.shape[0], 3), np.float32)
            s, c = self.model(torch.from_numpy(xyzs.reshape(-1, 3)[valid]), torch.from_numpy(dirs.reshape(-1, 3)[valid]))
            sig[valid] = s.numpy(); col[valid] = c.numpy()
            orc.composite_test_fw(sig.reshape(-1, ns), col.reshape(-1, ns, 3), deltas, ts, h, alive, T_thr, n_eff, opacity, depth, rgb)
            alive = alive[alive >= 0]
        bg = 1.0 if self.esf == 0 else 0.0
        return dict(rgb=rgb + bg * (1 - opacity)[:, None], depth=depth, opacity=opacity, total_samples=total)
