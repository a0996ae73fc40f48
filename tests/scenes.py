"""Seeded inputs shared by the oracle tests, the GPU parity tests and the golden-vector generator."""
import numpy as np

from mfnerf_b200 import synthetic as syn

NEAR = 0.01


def scene(name, n_rays, seed=0):
    """-> dict(rays_o, rays_d, bitfield, cascades, scale, esf, grid_size, max_samples, noise, center, half)"""
    if name == "lego":          # BASELINE config 1/2: scale 0.5, one cascade, constant step
        scale, cascades, esf = 0.5, 1, 0.0
        grid = syn.lego_density_grid(scale, cascades)
        o, d, _, _ = syn.random_rays(n_rays, seed)
    elif name == "full":        # warm-up phase: every cell occupied
        scale, cascades, esf = 0.5, 1, 0.0
        grid = np.ones((1, 128 ** 3), np.float32)
        o, d, _, _ = syn.random_rays(n_rays, seed)
    elif name == "unbounded":   # BASELINE config 4: scale 16, 6 cascades, exponential stepping
        scale, cascades, esf = 16.0, 6, 1.0 / 256
        grid = syn.lego_density_grid(scale, cascades)
        o, d, _, _ = syn.random_rays(n_rays, seed)
    elif name == "axis":        # axis-parallel rays: 1/0 = inf paths, some start inside the box
        scale, cascades, esf = 0.5, 1, 0.0
        grid = syn.lego_density_grid(scale, cascades)
        rng = np.random.RandomState(seed)
        o = rng.uniform(-0.45, 0.45, (n_rays, 3)).astype(np.float32)
        axis = rng.randint(0, 3, n_rays); sign = rng.choice([-1.0, 1.0], n_rays)
        d = np.zeros((n_rays, 3), np.float32); d[np.arange(n_rays), axis] = sign
        o[: n_rays // 2, 0] = -1.5 * np.sign(d[: n_rays // 2, 0] + 0.5)   # half of them start outside
    else:
        raise KeyError(name)
    bits = syn.bitfield_from_grid(grid)
    rng = np.random.RandomState(seed + 7)
    return dict(rays_o=o, rays_d=d, bitfield=bits, cascades=cascades, scale=scale, esf=esf, grid_size=128, max_samples=1024,
                noise=rng.rand(n_rays).astype(np.float32), center=np.zeros((1, 3), np.float32),
                half=np.full((1, 3), scale, np.float32))


def near_clamp(hits_t):
    """rendering.py:29 -- hits_t (R,1,2) -> (R,2) with 0<=t1<NEAR bumped to NEAR"""
    h = np.ascontiguousarray(hits_t[:, 0]).copy()
    m = (h[:, 0] >= 0) & (h[:, 0] < NEAR)
    h[m, 0] = NEAR
    return h


def field_values(n, seed=0):
    """seeded stand-ins for the network outputs: sigmas (n) >= 0 with a heavy tail, rgbs (n,3) in [0,1]"""
    rng = np.random.RandomState(seed + 11)
    sig = np.exp(rng.normal(1.0, 2.5, n)).astype(np.float32)
    rgb = rng.rand(n, 3).astype(np.float32)
    return sig, rgb
