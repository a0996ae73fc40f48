"""CPU-side checks of the drop-in boundary: the C-ABI library loads (no GPU needed for dlopen) and exports every
symbol include/mfnerf_b200.h declares; the vren drop-in exposes the reference's 12 names."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mfnerf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mfnerf_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mfn_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(_lib.FUNCS), declared ^ set(_lib.FUNCS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib.mfn_version() == 100
    assert _lib.lib.mfn_march_train_workspace_bytes(8192, 1024) == 256 + 8192 * 4 + 8192 * 1024 * 8    # [queue header | counts | stash]


def test_vren_dropin_surface():
    """ref: models/csrc/binding.cpp:234-251 registers exactly these 12 functions"""
    import vren
    names = ["ray_aabb_intersect", "ray_sphere_intersect", "morton3D", "morton3D_invert", "packbits", "raymarching_train", "raymarching_test",
             "composite_train_fw", "composite_train_bw", "composite_test_fw", "distortion_loss_fw", "distortion_loss_bw"]
    for n in names:
        assert callable(getattr(vren, n)), n


def test_argument_errors_without_gpu():
    """argument validation happens before any CUDA call, so it is testable on the CPU box"""
    from mfnerf_b200 import _lib
    rc = _lib.lib.mfn_morton3d(None, -1, None, None)
    assert rc == -2 and b"bad argument" in _lib.lib.mfn_last_error()
    rc = _lib.lib.mfn_packbits(None, 7, 0, 0.0, None, None)
    assert rc == 0  # empty input is a no-op


def test_render_and_field_entry_points_validate_arguments():
    """the device-side render wavefront and the fused field ops reject bad arguments with MFN_ERR_ARG before touching CUDA"""
    from mfnerf_b200 import _lib
    from mfnerf_b200.engine import make_field_cfg
    lib = _lib.lib
    assert lib.mfn_render_workspace_bytes(640000, 1) >= 640000 * (8 + 8 + 4 + 24)   # hits_t, 2 alive lists, N_eff + 24 B per sample row (t, dt, sigma, rgb)
    assert lib.mfn_render_workspace_bytes(-1, 1) == -1 and lib.mfn_render_workspace_bytes(10, 0) == -1
    assert lib.mfn_render_begin(None, None, None, None, 16, 0.01, 1, 1024, None, None, None, None, 0, None) == -2
    assert b"mfn_render_begin" in lib.mfn_last_error()
    cfg = make_field_cfg(0.5)
    assert lib.mfn_render_iterations(ctypes.byref(cfg), None, None, None, None, 16, None, 1, 0.5, 0.0, 128, 1024, 1, 1e-4, 4, None, None, None, None, 0, None) == -2
    wide = make_field_cfg(0.5, rgb_channels=128)          # 128-wide rgb net (MF-NeRF's scripts): on the fused kernels since round 2
    assert lib.mfn_field_is_fused(ctypes.byref(wide)) == 1 and lib.mfn_field_is_fused(ctypes.byref(cfg)) == 1
    odd = make_field_cfg(0.5, rgb_channels=96)            # anything else: loud error, no silent fallback
    assert lib.mfn_field_is_fused(ctypes.byref(odd)) == -2 and b"rgb net" in lib.mfn_last_error()
    assert lib.mfn_render_finish(None, None, None, 16, None) == -2
    assert lib.mfn_render_iterations(ctypes.byref(cfg), None, None, None, None, 0, None, 1, 0.5, 0.0, 128, 1024, 1, 1e-4, 4, None, None, None, None, 0, None) == 0   # no rays: no-op
    assert lib.mfn_field_workspace_bytes(ctypes.byref(cfg), 1 << 20, 1) >= (1 << 20) * (64 + 8 + 64 + 16 + 12)   # X tile, fp16 rgb, dfeats, x01, dirs per sample
    assert lib.mfn_field_fwd(ctypes.byref(cfg), None, None, None, None, 128, None, None, None, None, 0, None) == -2
    assert lib.mfn_geo_fwd(ctypes.byref(cfg), None, None, 128, None, None, None) == -2 and b"null pointer" in lib.mfn_last_error()
    assert lib.mfn_geo_fwd(ctypes.byref(cfg), None, None, 0, None, None, None) == 0
    # MixedFeature grid: K tables of 2^T entries; fused like the plain hash grid
    from mfnerf_b200 import field_ops
    mixed = make_field_cfg(0.5, log2_T=17, grid="MixedFeature", n_tables=8, rgb_channels=128)
    assert field_ops.grid_layout(mixed.grid)[0] == 8 << 17 and lib.mfn_field_is_fused(ctypes.byref(mixed)) == 1
    assert lib.mfn_grid_layout(ctypes.byref(field_ops.make_grid_cfg(16, 2, 17, 16, 1.3, "MixedFeature", 0)), None, None, None) == -1
    with pytest.raises(NotImplementedError):
        field_ops.make_grid_cfg(16, 2, 17, 16, 1.3, "Window", 1)
    # ray generation / mark_invisible_cells
    from mfnerf_b200.dataset import Camera
    cam = Camera(100.0, 100.0, 32.0, 24.0, 64, 48)
    rb = lambda cam, n_img, n, draw=0, image=0: lib.mfn_ray_batch(ctypes.byref(cam), None, None, n_img, None, 0, None, None, image, draw, 1, None, n, None, None, None, None, None, None)
    assert rb(cam, 4, 0) == 0                                               # no rays: no-op
    assert rb(cam, 4, 16) == -2 and b"null pointer" in lib.mfn_last_error()
    assert rb(cam, 0, 16) == -2 and rb(cam, 4, 16, draw=3) == -2 and rb(Camera(0.0, 1.0, 0.0, 0.0, 8, 8), 4, 16) == -2
    assert lib.mfn_grid_mark_invisible(None, None, 4, 64, 48, 1, 0.5, 128, 0.01, None, None, None) == -2
    assert lib.mfn_grid_mark_invisible(None, None, 0, 64, 48, 1, 0.5, 128, 0.01, None, None, None) == -2 and b"bad argument" in lib.mfn_last_error()


def test_host_side_helpers_and_new_entry_points_validate_arguments():
    """mfn_adam_hyper is pure host code; the fused step entry points reject null pointers before touching CUDA"""
    import numpy as np
    from mfnerf_b200 import _lib
    from mfnerf_b200.engine import make_field_cfg
    lib = _lib.lib
    h = (ctypes.c_float * 4)()
    for step in (1, 2, 17, 1000, 30000):
        assert lib.mfn_adam_hyper(1e-2, 0.9, 0.999, step, h) == 0
        assert h[0] == np.float32(1e-2)
        b1, b2 = float(np.float32(0.9)), float(np.float32(0.999))                                     # the betas are floats; the powers are taken in double (apex)
        assert h[1] == np.float32(1 - b1 ** step) and h[2] == np.float32(1 - b2 ** step)
    assert lib.mfn_adam_hyper(1e-2, 0.9, 0.999, 0, h) == -2 and lib.mfn_adam_hyper(1e-2, 0.9, 0.999, 1, None) == -2
    assert lib.mfn_adam_step_dev(None, None, None, None, None, 16, None, 0.9, 0.999, 1e-15, 1.0, None, 1, None) == -2
    assert lib.mfn_adam_step(None, None, None, None, None, 0, 1e-2, 0.9, 0.999, 1e-15, 1, 1.0, None, 1, None) == 0          # no parameters: no-op
    bg = (ctypes.c_float * 3)(1.0, 1.0, 1.0)
    args = [None] * 6 + [1e-4, 16, 100, bg, 1e-3, 1.0] + [None] * 13 + [None]
    assert lib.mfn_composite_loss_train(*args) == -2 and b"mfn_composite_loss_train" in lib.mfn_last_error()
    args[7] = 0
    assert lib.mfn_composite_loss_train(*args) == 0                                                                       # no rays: no-op
    cfg = make_field_cfg(0.5)
    assert lib.mfn_field_count_ptr(ctypes.byref(cfg), None, 1024) is None
    fake = ctypes.c_void_p(1 << 20)                                                                                         # only address arithmetic happens
    p = lib.mfn_field_count_ptr(ctypes.byref(cfg), fake, 1024)
    assert p is not None and (1 << 20) < p < (1 << 20) + lib.mfn_field_workspace_bytes(ctypes.byref(cfg), 1024, 1)
    wide = make_field_cfg(0.5, rgb_channels=128)
    pw = lib.mfn_field_count_ptr(ctypes.byref(wide), fake, 1024)
    assert pw is not None and pw > p                                                                                       # larger weight-gradient partials
