"""Drop-in replacement for the reference's pybind11 module `vren` (models/csrc/binding.cpp:234-251).

Same 12 names, argument order, return tuples, dtypes, shapes and in-place mutations; the work is done by
hand-written sm_100a kernels in libmfnerf_b200.so, reached through ctypes.  Tensors are allocated here with
torch and handed over as raw device pointers on torch's *current* stream (the reference always used the
legacy default stream).

Error convention (ref: include/utils.h:4-6): RuntimeError "<name> must be a CUDA tensor" / "<name> must be
contiguous", plus dtype and shape checks the reference left to crash inside its accessors.

One documented deviation: raymarching_train returns xyzs/dirs/deltas/ts with exactly `counter[0]` rows
instead of N_rays*max_samples zero-filled rows (the reference slices them to counter[0] immediately,
custom_functions.py:91-96), which removes a 268 MB memset per step.
"""
import ctypes

import torch

from mfnerf_b200 import _lib
from mfnerf_b200._lib import call, ptr, stream_ptr

__all__ = [
    "ray_aabb_intersect", "ray_sphere_intersect", "morton3D", "morton3D_invert", "packbits", "raymarching_train",
    "raymarching_test", "composite_train_fw", "composite_train_bw", "composite_test_fw", "distortion_loss_fw",
    "distortion_loss_bw",
]

_DT = {torch.float32: 0, torch.float16: 1, torch.float64: 2}


def _chk(t, name, dtype=None, ndim=None, last=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    if last is not None and t.shape[-1] != last:
        raise RuntimeError(f"{name} must have last dim {last}, got shape {tuple(t.shape)}")
    return t


def _f32(t, name, ndim=None, last=None):
    return _chk(t, name, torch.float32, ndim, last)


# ------------------------------------------------------------------------------------------------ intersection
def _intersect(fn, rays_o, rays_d, centers, extent, max_hits, extent_ndim):
    _f32(rays_o, "rays_o", 2, 3); _f32(rays_d, "rays_d", 2, 3); _f32(centers, "centers", 2, 3)
    _f32(extent, "half_sizes" if extent_ndim == 2 else "radii", extent_ndim)
    n, v = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    with torch.cuda.device(dev):
        hit_cnt = torch.empty(n, dtype=torch.int32, device=dev)
        hits_t = torch.empty(n, max_hits, 2, dtype=torch.float32, device=dev)
        hits_idx = torch.empty(n, max_hits, dtype=torch.int64, device=dev)
        call(fn, ptr(rays_o), ptr(rays_d), ptr(centers), ptr(extent), n, v, int(max_hits), ptr(hit_cnt), ptr(hits_t),
             ptr(hits_idx), stream_ptr(dev))
    return [hit_cnt, hits_t, hits_idx]


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """ref: binding.cpp:4-16.  -> [hit_cnt (N) i32, hits_t (N,max_hits,2) f32, hits_voxel_idx (N,max_hits) i64]"""
    return _intersect("mfn_ray_aabb_intersect", rays_o, rays_d, centers, half_sizes, max_hits, 2)


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    """ref: binding.cpp:19-32."""
    return _intersect("mfn_ray_sphere_intersect", rays_o, rays_d, centers, radii, max_hits, 1)


# ------------------------------------------------------------------------------------------------ grid utils
def morton3D(coords):
    """ref: binding.cpp:47-51.  coords (N,3) int32 -> indices (N) int32"""
    _chk(coords, "coords", torch.int32, 2, 3)
    out = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
    with torch.cuda.device(coords.device):
        call("mfn_morton3d", ptr(coords), coords.shape[0], ptr(out), stream_ptr(coords.device))
    return out


def morton3D_invert(indices):
    """ref: binding.cpp:54-58.  indices (N) int32 -> coords (N,3) int32"""
    _chk(indices, "indices", torch.int32, 1)
    out = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=indices.device)
    with torch.cuda.device(indices.device):
        call("mfn_morton3d_invert", ptr(indices), indices.shape[0], ptr(out), stream_ptr(indices.device))
    return out


def packbits(density_grid, density_threshold, density_bitfield):
    """ref: binding.cpp:35-44.  writes density_bitfield (uint8) in place; returns None"""
    _chk(density_grid, "density_grid"); _chk(density_bitfield, "density_bitfield", torch.uint8)
    if density_grid.dtype not in _DT:
        raise RuntimeError(f"density_grid must be float16/32/64, got {density_grid.dtype}")
    n_bytes = density_bitfield.shape[0]
    if density_grid.numel() < 8 * n_bytes:
        raise RuntimeError("density_grid has fewer than 8*len(density_bitfield) cells")
    with torch.cuda.device(density_grid.device):
        call("mfn_packbits", ptr(density_grid), _DT[density_grid.dtype], n_bytes, float(density_threshold),
             ptr(density_bitfield), stream_ptr(density_grid.device))
    return None


# ------------------------------------------------------------------------------------------------ marching
_ws_cache = {}


def _workspace(dev, nbytes):
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size,
                      max_samples):
    """ref: binding.cpp:60-81.  -> [rays_a (N,3) i64, xyzs (S,3), dirs (S,3), deltas (S), ts (S), counter (2) i32]
    with S = counter[0] (see module docstring)."""
    _f32(rays_o, "rays_o", 2, 3); _f32(rays_d, "rays_d", 2, 3); _f32(hits_t, "hits_t", 2, 2)
    _chk(density_bitfield, "density_bitfield", torch.uint8, 1); _f32(noise, "noise", 1)
    n = rays_o.shape[0]
    if density_bitfield.numel() * 8 < int(cascades) * int(grid_size) ** 3:
        raise RuntimeError("density_bitfield is smaller than cascades*grid_size**3/8 bytes")
    dev = rays_o.device
    with torch.cuda.device(dev):
        st = stream_ptr(dev)
        ws_bytes = _lib.lib.mfn_march_train_workspace_bytes(n, int(max_samples))
        ws = _workspace(dev, ws_bytes)
        rays_a = torch.empty(n, 3, dtype=torch.int64, device=dev)
        counter = torch.empty(2, dtype=torch.int32, device=dev)
        call("mfn_march_train_count", ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(density_bitfield), int(cascades), float(scale),
             float(exp_step_factor), ptr(noise), int(grid_size), int(max_samples), n, ptr(rays_a), ptr(counter), ptr(ws),
             ws.numel(), st)
        total = int(counter[0].item())  # the reference syncs at the same place (custom_functions.py:91-96)
        xyzs = torch.empty(total, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(total, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(total, dtype=torch.float32, device=dev)
        ts = torch.empty(total, dtype=torch.float32, device=dev)
        call("mfn_march_train_write", ptr(rays_o), ptr(rays_d), ptr(rays_a), ptr(ws), int(max_samples), n, total, ptr(xyzs),
             ptr(dirs), ptr(deltas), ptr(ts), st)
    return [rays_a, xyzs, dirs, deltas, ts, counter]


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                     max_samples, N_samples):
    """ref: binding.cpp:84-107.  hits_t (N,2) is advanced in place.
    -> [xyzs (A,Ns,3), dirs (A,Ns,3), deltas (A,Ns), ts (A,Ns), N_eff_samples (A) i32]"""
    _f32(rays_o, "rays_o", 2, 3); _f32(rays_d, "rays_d", 2, 3); _f32(hits_t, "hits_t", 2, 2)
    _chk(alive_indices, "alive_indices", torch.int64, 1); _chk(density_bitfield, "density_bitfield", torch.uint8, 1)
    a, ns = alive_indices.shape[0], int(N_samples)
    dev = rays_o.device
    with torch.cuda.device(dev):
        xyzs = torch.empty(a, ns, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(a, ns, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(a, ns, dtype=torch.float32, device=dev)
        ts = torch.empty(a, ns, dtype=torch.float32, device=dev)
        n_eff = torch.empty(a, dtype=torch.int32, device=dev)
        call("mfn_raymarching_test", ptr(rays_o), ptr(rays_d), ptr(hits_t), ptr(alive_indices), ptr(density_bitfield),
             int(cascades), float(scale), float(exp_step_factor), int(grid_size), int(max_samples), ns, a, ptr(xyzs), ptr(dirs),
             ptr(deltas), ptr(ts), ptr(n_eff), stream_ptr(dev))
    return [xyzs, dirs, deltas, ts, n_eff]


# ------------------------------------------------------------------------------------------------ compositing
def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, opacity_threshold):
    """ref: binding.cpp:110-127.  -> [total_samples (R) i64, opacity (R), depth (R), rgb (R,3), ws (N)]"""
    _f32(sigmas, "sigmas", 1); _f32(rgbs, "rgbs", 2, 3); _f32(deltas, "deltas", 1); _f32(ts, "ts", 1)
    _chk(rays_a, "rays_a", torch.int64, 2, 3)
    r, n = rays_a.shape[0], sigmas.shape[0]
    dev = sigmas.device
    with torch.cuda.device(dev):
        total = torch.zeros(r, dtype=torch.int64, device=dev)
        opacity = torch.zeros(r, dtype=torch.float32, device=dev)
        depth = torch.zeros(r, dtype=torch.float32, device=dev)
        rgb = torch.zeros(r, 3, dtype=torch.float32, device=dev)
        ws = torch.empty(n, dtype=torch.float32, device=dev)
        call("mfn_composite_train_fw", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(ts), ptr(rays_a), float(opacity_threshold), r, n,
             ptr(total), ptr(opacity), ptr(depth), ptr(rgb), ptr(ws), stream_ptr(dev))
    return [total, opacity, depth, rgb, ws]


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb,
                       opacity_threshold):
    """ref: binding.cpp:130-167.  -> [dL_dsigmas (N), dL_drgbs (N,3)]"""
    _f32(dL_dopacity, "dL_dopacity", 1); _f32(dL_ddepth, "dL_ddepth", 1); _f32(dL_drgb, "dL_drgb", 2, 3); _f32(dL_dws, "dL_dws", 1)
    _f32(sigmas, "sigmas", 1); _f32(rgbs, "rgbs", 2, 3); _f32(ws, "ws", 1); _f32(deltas, "deltas", 1); _f32(ts, "ts", 1)
    _chk(rays_a, "rays_a", torch.int64, 2, 3); _f32(opacity, "opacity", 1); _f32(depth, "depth", 1); _f32(rgb, "rgb", 2, 3)
    r, n = rays_a.shape[0], sigmas.shape[0]
    dev = sigmas.device
    with torch.cuda.device(dev):
        dL_dsigmas = torch.empty(n, dtype=torch.float32, device=dev)
        dL_drgbs = torch.empty(n, 3, dtype=torch.float32, device=dev)
        call("mfn_composite_train_bw", ptr(dL_dopacity), ptr(dL_ddepth), ptr(dL_drgb), ptr(dL_dws), ptr(sigmas), ptr(rgbs), ptr(ws),
             ptr(deltas), ptr(ts), ptr(rays_a), ptr(opacity), ptr(depth), ptr(rgb), float(opacity_threshold), r, n,
             ptr(dL_dsigmas), ptr(dL_drgbs), stream_ptr(dev))
    return [dL_dsigmas, dL_drgbs]


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples, opacity, depth, rgb):
    """ref: binding.cpp:170-198.  Updates opacity/depth/rgb and alive_indices in place; returns None.
    (hits_t is accepted and ignored, exactly like the reference kernel.)"""
    _f32(sigmas, "sigmas", 2); _f32(rgbs, "rgbs", 3, 3); _f32(deltas, "deltas", 2); _f32(ts, "ts", 2); _chk(hits_t, "hits_t")
    _chk(alive_indices, "alive_indices", torch.int64, 1); _chk(N_eff_samples, "N_eff_samples", torch.int32, 1)
    _f32(opacity, "opacity", 1); _f32(depth, "depth", 1); _f32(rgb, "rgb", 2, 3)
    a, ns = alive_indices.shape[0], sigmas.shape[1]
    dev = sigmas.device
    with torch.cuda.device(dev):
        call("mfn_composite_test_fw", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(ts), ptr(alive_indices), float(T_threshold),
             ptr(N_eff_samples), ns, a, ptr(opacity), ptr(depth), ptr(rgb), stream_ptr(dev))
    return None


# ------------------------------------------------------------------------------------------------ distortion loss
def distortion_loss_fw(ws, deltas, ts, rays_a):
    """ref: binding.cpp:201-213.  -> [loss (R), ws_inclusive_scan (N), wts_inclusive_scan (N)]"""
    _f32(ws, "ws", 1); _f32(deltas, "deltas", 1); _f32(ts, "ts", 1); _chk(rays_a, "rays_a", torch.int64, 2, 3)
    r, n = rays_a.shape[0], ws.shape[0]
    dev = ws.device
    with torch.cuda.device(dev):
        loss = torch.zeros(r, dtype=torch.float32, device=dev)
        wi = torch.empty(n, dtype=torch.float32, device=dev)
        wti = torch.empty(n, dtype=torch.float32, device=dev)
        call("mfn_distortion_loss_fw", ptr(ws), ptr(deltas), ptr(ts), ptr(rays_a), r, n, ptr(loss), ptr(wi), ptr(wti),
             stream_ptr(dev))
    return [loss, wi, wti]


def distortion_loss_bw(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a):
    """ref: binding.cpp:216-231.  -> dL_dws (N)"""
    _f32(dL_dloss, "dL_dloss", 1); _f32(ws_inclusive_scan, "ws_inclusive_scan", 1); _f32(wts_inclusive_scan, "wts_inclusive_scan", 1)
    _f32(ws, "ws", 1); _f32(deltas, "deltas", 1); _f32(ts, "ts", 1); _chk(rays_a, "rays_a", torch.int64, 2, 3)
    r, n = rays_a.shape[0], ws.shape[0]
    dev = ws.device
    with torch.cuda.device(dev):
        out = torch.empty(n, dtype=torch.float32, device=dev)
        call("mfn_distortion_loss_bw", ptr(dL_dloss), ptr(ws_inclusive_scan), ptr(wts_inclusive_scan), ptr(ws), ptr(deltas), ptr(ts),
             ptr(rays_a), r, n, ptr(out), stream_ptr(dev))
    return out
