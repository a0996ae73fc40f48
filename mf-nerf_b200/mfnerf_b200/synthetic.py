"""Seeded synthetic "Lego-shaped" scene used by the tests and bench.py (no datasets: there is no network).

Cameras follow the reference's Blender reader (datasets/nerf.py:26-27,70-72: 800x800, fx = 0.5*800/tan(0.5*0.6911112),
positions scaled to radius 1.5) and ray generation follows datasets/ray_utils.py:34-35,60-68 (pixel-centre,
un-normalised directions, rays_d = dir @ R^T, rays_o = t).  The occupancy grid is a union of axis-aligned
boxes (tracks, chassis, cab, arm, blade) voxelised at 128^3 in morton order, ~5 % occupied.
"""
import math

import numpy as np
import torch

IMG_WH = (800, 800)
FOCAL = 0.5 * 800 / math.tan(0.5 * 0.6911112)
CAM_RADIUS = 1.5

# (xmin, ymin, zmin, xmax, ymax, zmax, r, g, b) in world units, all inside [-0.35, 0.35]^3
LEGO_BOXES = np.array([
    [-0.30, -0.22, -0.30, 0.30, -0.12, -0.18, 0.20, 0.20, 0.20],   # left track
    [-0.30, 0.12, -0.30, 0.30, 0.22, -0.18, 0.20, 0.20, 0.20],     # right track
    [-0.26, -0.14, -0.20, 0.24, 0.14, -0.08, 0.85, 0.70, 0.10],    # chassis
    [-0.22, -0.12, -0.08, -0.02, 0.12, 0.12, 0.90, 0.75, 0.12],    # cab
    [-0.20, -0.10, 0.12, -0.04, 0.10, 0.16, 0.80, 0.65, 0.10],     # cab roof
    [-0.02, -0.10, -0.08, 0.22, 0.10, 0.00, 0.85, 0.70, 0.10],     # engine hood
    [0.02, -0.03, 0.00, 0.08, 0.03, 0.22, 0.60, 0.60, 0.62],       # arm riser
    [0.06, -0.03, 0.18, 0.30, 0.03, 0.24, 0.60, 0.60, 0.62],       # arm boom
    [0.26, -0.03, 0.02, 0.32, 0.03, 0.20, 0.60, 0.60, 0.62],       # arm stick
    [0.24, -0.16, -0.30, 0.34, 0.16, -0.12, 0.75, 0.75, 0.78],     # blade
    [-0.34, -0.08, -0.16, -0.26, 0.08, -0.04, 0.30, 0.30, 0.32],   # rear weight
    [-0.12, -0.18, -0.12, 0.10, -0.14, -0.02, 0.85, 0.20, 0.15],   # side panel
], dtype=np.float32)


def morton_order_coords(G=128):
    """(G^3, 3) int32 cell coords listed in morton order: row m holds morton3D_invert(m)."""
    m = np.arange(G ** 3, dtype=np.uint32)

    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xC30C30C3
        x = (x | (x >> 4)) & 0x0F00F00F
        x = (x | (x >> 8)) & 0xFF0000FF
        x = (x | (x >> 16)) & 0x0000FFFF
        return x
    return np.stack([compact(m), compact(m >> 1), compact(m >> 2)], 1).astype(np.int32)


def lego_density_grid(scale=0.5, cascades=1, G=128, inside=10.0):
    """(cascades, G^3) float32 density grid in morton order; cells whose centre is inside a box get `inside`."""
    coords = morton_order_coords(G).astype(np.float32)
    grid = np.zeros((cascades, G ** 3), np.float32)
    for c in range(cascades):
        s = min(2.0 ** (c - 1), scale)
        centre = ((coords + 0.5) / G * 2 - 1) * s
        occ = np.zeros(G ** 3, bool)
        for b in LEGO_BOXES:
            occ |= np.all((centre >= b[0:3]) & (centre <= b[3:6]), axis=1)
        if c > 0:  # sparse "ground slab" in the outer cascades of unbounded scenes
            occ |= (np.abs(centre[:, 2] + 0.32) < s / G) & (np.abs(centre[:, 0]) < 0.8 * s) & (np.abs(centre[:, 1]) < 0.8 * s)
        grid[c, occ] = inside
    return grid


def bitfield_from_grid(grid, thr=0.5):
    """numpy packbits, little bit order == bit i of byte n is cell 8n+i (ref: raymarching.cu:136-138)."""
    return np.packbits((grid.reshape(-1) > thr), bitorder="little")


def camera_poses(n, seed=0, radius=CAM_RADIUS):
    """(n, 3, 4) float32 camera-to-world matrices on the upper hemisphere looking at the origin
    (camera convention right-down-front as in datasets/ray_utils.py)."""
    rng = np.random.RandomState(seed)
    theta = rng.uniform(0, 2 * np.pi, n)
    phi = rng.uniform(np.deg2rad(10), np.deg2rad(80), n)  # elevation
    pos = radius * np.stack([np.cos(phi) * np.cos(theta), np.cos(phi) * np.sin(theta), np.sin(phi)], 1)
    fwd = -pos / np.linalg.norm(pos, axis=1, keepdims=True)
    up = np.array([0, 0, 1.0])
    right = np.cross(fwd, up); right /= np.linalg.norm(right, axis=1, keepdims=True)
    down = np.cross(fwd, right)
    c2w = np.stack([right, down, fwd, pos], 2)  # columns: x=right, y=down, z=front, t
    return c2w.astype(np.float32)


def pixel_directions(pix_idx, wh=IMG_WH, focal=FOCAL):
    """pixel-centre camera-space directions ((u-cx+.5)/fx, (v-cy+.5)/fy, 1), un-normalised (ray_utils.py:34-35)"""
    w, h = wh
    u = (pix_idx % w).astype(np.float32); v = (pix_idx // w).astype(np.float32)
    return np.stack([(u - w / 2 + 0.5) / focal, (v - h / 2 + 0.5) / focal, np.ones_like(u)], 1).astype(np.float32)


def random_rays(n, seed=0, n_cams=100, wh=IMG_WH):
    """n training rays sampled like ray_sampling_strategy='all_images' (datasets/base.py:26-30).
    -> rays_o (n,3), rays_d (n,3) float32 numpy, plus (img_idx, pix_idx)."""
    rng = np.random.RandomState(seed + 1)
    poses = camera_poses(n_cams, seed)
    img = rng.randint(0, n_cams, n)
    pix = rng.randint(0, wh[0] * wh[1], n)
    d_cam = pixel_directions(pix, wh)
    R = poses[img, :, :3]
    rays_d = np.einsum("nij,nj->ni", R, d_cam).astype(np.float32)   # dir @ R^T
    rays_o = poses[img, :, 3].copy()
    return rays_o, rays_d, img, pix


def image_rays(pose, wh=IMG_WH):
    """all wh[0]*wh[1] rays of one view, row-major pixels"""
    pix = np.arange(wh[0] * wh[1])
    d_cam = pixel_directions(pix, wh)
    rays_d = (d_cam @ pose[:, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, rays_d


def analytic_render(rays_o, rays_d, boxes=LEGO_BOXES, bg=1.0):
    """Ground-truth colours for the synthetic scene: nearest box hit -> that box's colour shaded by the hit
    face, else white background.  Gives the training loop a learnable target without any dataset."""
    o = torch.as_tensor(rays_o, dtype=torch.float32); d = torch.as_tensor(rays_d, dtype=torch.float32)
    bx = torch.as_tensor(boxes, dtype=torch.float32, device=o.device)
    inv = 1.0 / d
    t_best = torch.full((o.shape[0],), float("inf"), device=o.device)
    col = torch.full((o.shape[0], 3), bg, device=o.device)
    for b in bx:
        tmin = (b[0:3] - o) * inv; tmax = (b[3:6] - o) * inv
        t1 = torch.minimum(tmin, tmax); t2 = torch.maximum(tmin, tmax)
        tn, axis = t1.max(1); tf = t2.min(1).values
        hit = (tn <= tf) & (tf > 0) & (tn < t_best) & (tn > 0)
        shade = 0.6 + 0.2 * axis.float()
        c = b[6:9][None] * shade[:, None]
        col = torch.where(hit[:, None], c, col); t_best = torch.where(hit, tn, t_best)
    return col.clamp(0, 1)
