"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of oracle/vren_oracle.c (CPU restatement of the
reference's vren kernels, models/csrc/*.cu).  Same function names and return tuples as the reference's
`vren` module (binding.cpp:234-251) but on numpy arrays.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product (mf-nerf_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_vren.so")


def build(force=False):
    src = os.path.join(_HERE, "vren_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_vren.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_raymarching_train.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def morton3D(coords):
    coords = _c(coords, np.int32)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3d(_p(coords), ctypes.c_int64(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _c(indices, np.int32)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3d_invert(_p(indices), ctypes.c_int64(indices.shape[0]), _p(out))
    return out


def packbits(density_grid, thr, density_bitfield):
    g = _c(density_grid, np.float32).reshape(-1)
    lib().orc_packbits_f32(_p(g), ctypes.c_int64(density_bitfield.shape[0]), ctypes.c_float(thr), _p(density_bitfield))


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    o, d, c, h = (_c(x, np.float32) for x in (rays_o, rays_d, centers, half_sizes))
    n, v = o.shape[0], c.shape[0]
    cnt = np.empty(n, np.int32); ht = np.empty((n, max_hits, 2), np.float32); hi = np.empty((n, max_hits), np.int64)
    lib().orc_ray_aabb_intersect(_p(o), _p(d), _p(c), _p(h), ctypes.c_int64(n), ctypes.c_int64(v), ctypes.c_int(max_hits), _p(cnt), _p(ht), _p(hi))
    return cnt, ht, hi


def raymarching_train(rays_o, rays_d, hits_t, bitfield, cascades, scale, esf, noise, grid_size, max_samples):
    """-> rays_a (R,3) i64 [ray_idx, start (prefix sum), N], xyzs, dirs, deltas, ts (exactly total rows), counter"""
    o, d, h, nz = (_c(x, np.float32) for x in (rays_o, rays_d, hits_t, noise))
    bits = _c(bitfield, np.uint8)
    n = o.shape[0]
    rays_a = np.empty((n, 3), np.int64)
    # first call with capacity 0 to learn the total, then the real one
    dummy = np.empty(1, np.float32)
    args = lambda cap, x, dd, dl, t: (_p(o), _p(d), _p(h), _p(bits), ctypes.c_int(cascades), ctypes.c_float(scale), ctypes.c_float(esf), _p(nz),
                                      ctypes.c_int(grid_size), ctypes.c_int(max_samples), ctypes.c_int64(n), ctypes.c_int64(cap), _p(rays_a),
                                      _p(x), _p(dd), _p(dl), _p(t))
    total = lib().orc_raymarching_train(*args(0, dummy, dummy, dummy, dummy))
    xyzs = np.empty((total, 3), np.float32); dirs = np.empty((total, 3), np.float32)
    deltas = np.empty(total, np.float32); ts = np.empty(total, np.float32)
    lib().orc_raymarching_train(*args(total, xyzs, dirs, deltas, ts))
    return rays_a, xyzs, dirs, deltas, ts, np.array([total, n], np.int32)


def raymarching_test(rays_o, rays_d, hits_t, alive, bitfield, cascades, scale, esf, grid_size, max_samples, n_samples):
    """hits_t (R,2) float32 numpy array is advanced in place"""
    o, d = _c(rays_o, np.float32), _c(rays_d, np.float32)
    assert hits_t.dtype == np.float32 and hits_t.flags.c_contiguous
    alive = _c(alive, np.int64); bits = _c(bitfield, np.uint8)
    a = alive.shape[0]
    xyzs = np.empty((a, n_samples, 3), np.float32); dirs = np.empty((a, n_samples, 3), np.float32)
    deltas = np.empty((a, n_samples), np.float32); ts = np.empty((a, n_samples), np.float32); n_eff = np.empty(a, np.int32)
    lib().orc_raymarching_test(_p(o), _p(d), _p(hits_t), _p(alive), _p(bits), ctypes.c_int(cascades), ctypes.c_float(scale), ctypes.c_float(esf),
                               ctypes.c_int(grid_size), ctypes.c_int(max_samples), ctypes.c_int(n_samples), ctypes.c_int64(a), _p(xyzs), _p(dirs),
                               _p(deltas), _p(ts), _p(n_eff))
    return xyzs, dirs, deltas, ts, n_eff


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_thr):
    s, c, dl, t = (_c(x, np.float32) for x in (sigmas, rgbs, deltas, ts)); ra = _c(rays_a, np.int64)
    r, n = ra.shape[0], s.shape[0]
    total = np.zeros(r, np.int64); op = np.zeros(r, np.float32); dp = np.zeros(r, np.float32); rgb = np.zeros((r, 3), np.float32)
    ws = np.zeros(n, np.float32)
    lib().orc_composite_train_fw(_p(s), _p(c), _p(dl), _p(t), _p(ra), ctypes.c_float(T_thr), ctypes.c_int64(r), _p(total), _p(op), _p(dp), _p(rgb), _p(ws))
    return total, op, dp, rgb, ws


def composite_train_bw(gO, gD, gRGB, gW, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, T_thr):
    f = lambda x: _c(x, np.float32)
    gO, gD, gRGB, gW, sigmas, rgbs, ws, deltas, ts, opacity, depth, rgb = map(f, (gO, gD, gRGB, gW, sigmas, rgbs, ws, deltas, ts, opacity, depth, rgb))
    ra = _c(rays_a, np.int64)
    r, n = ra.shape[0], sigmas.shape[0]
    ds = np.zeros(n, np.float32); dc = np.zeros((n, 3), np.float32)
    lib().orc_composite_train_bw(_p(gO), _p(gD), _p(gRGB), _p(gW), _p(sigmas), _p(rgbs), _p(ws), _p(deltas), _p(ts), _p(ra), _p(opacity), _p(depth),
                                 _p(rgb), ctypes.c_float(T_thr), ctypes.c_int64(r), ctypes.c_int64(n), _p(ds), _p(dc))
    return ds, dc


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive, T_thr, n_eff, opacity, depth, rgb):
    """alive (int64), opacity, depth, rgb (float32, contiguous) are updated in place"""
    s, c, dl, t = (_c(x, np.float32) for x in (sigmas, rgbs, deltas, ts)); ne = _c(n_eff, np.int32)
    for x in (opacity, depth, rgb):
        assert x.dtype == np.float32 and x.flags.c_contiguous
    assert alive.dtype == np.int64 and alive.flags.c_contiguous
    lib().orc_composite_test_fw(_p(s), _p(c), _p(dl), _p(t), _p(alive), ctypes.c_float(T_thr), _p(ne), ctypes.c_int(s.shape[1]), ctypes.c_int64(alive.shape[0]),
                                _p(opacity), _p(depth), _p(rgb))


def distortion_loss_fw(ws, deltas, ts, rays_a):
    w, dl, t = (_c(x, np.float32) for x in (ws, deltas, ts)); ra = _c(rays_a, np.int64)
    r, n = ra.shape[0], w.shape[0]
    loss = np.zeros(r, np.float32); wi = np.zeros(n, np.float32); wti = np.zeros(n, np.float32)
    lib().orc_distortion_fw(_p(w), _p(dl), _p(t), _p(ra), ctypes.c_int64(r), _p(loss), _p(wi), _p(wti))
    return loss, wi, wti


def distortion_loss_bw(gL, ws_incl, wts_incl, ws, deltas, ts, rays_a):
    gL, wi, wti, w, dl, t = (_c(x, np.float32) for x in (gL, ws_incl, wts_incl, ws, deltas, ts)); ra = _c(rays_a, np.int64)
    out = np.zeros(w.shape[0], np.float32)
    lib().orc_distortion_bw(_p(gL), _p(wi), _p(wti), _p(w), _p(dl), _p(t), _p(ra), ctypes.c_int64(ra.shape[0]), _p(out))
    return out
