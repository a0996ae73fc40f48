"""Thin torch-facing wrappers of the field kernels in libmfnerf_b200.so (encoder / SH / fused MLPs / optimiser).
Tensors are allocated with torch; the kernels get raw pointers on torch's current stream."""
import ctypes

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

ACT = {"None": 0, "none": 0, None: 0, "Sigmoid": 1, "sigmoid": 1, "Exponential": 2, "exp": 2}
GRID_TYPES = {"Hash": 0, "MixedFeature": 1}      # MFN_GRID_HASH, MFN_GRID_MIXED ("Window": semantics unknown, not implemented)


class GridCfg(ctypes.Structure):
    """mirror of `mfn_grid_cfg` (include/mfnerf_b200.h)"""
    _fields_ = [("n_levels", ctypes.c_int32), ("n_features", ctypes.c_int32), ("log2_hashmap_size", ctypes.c_int32),
                ("base_resolution", ctypes.c_int32), ("per_level_scale", ctypes.c_double), ("grid_type", ctypes.c_int32),
                ("n_tables", ctypes.c_int32)]


class FieldCfg(ctypes.Structure):
    """mirror of `mfn_field_cfg` (include/mfnerf_b200.h)"""
    _fields_ = [("grid", GridCfg), ("sigma_width", ctypes.c_int32), ("sigma_hidden", ctypes.c_int32),
                ("rgb_width", ctypes.c_int32), ("rgb_hidden", ctypes.c_int32), ("rgb_act", ctypes.c_int32),
                ("xyz_min", ctypes.c_float * 3), ("xyz_max", ctypes.c_float * 3)]


def make_grid_cfg(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, grid_type="Hash", n_tables=1):
    if grid_type not in GRID_TYPES:
        raise NotImplementedError(f"grid type {grid_type!r} is not implemented (available: {sorted(GRID_TYPES)})")
    return GridCfg(int(n_levels), int(n_features), int(log2_hashmap_size), int(base_resolution), float(per_level_scale),
                   GRID_TYPES[grid_type], int(n_tables))


def grid_layout(cfg):
    """-> (total_entries, offsets[L+1], resolutions[L], scales[L]) as python lists"""
    L = cfg.n_levels
    off = (ctypes.c_uint32 * (L + 1))(); res = (ctypes.c_uint32 * L)(); sc = (ctypes.c_float * L)()
    total = _lib.lib.mfn_grid_layout(ctypes.byref(cfg), off, res, sc)
    if total < 0:
        raise _lib.MfnError(_lib.lib.mfn_last_error().decode())
    return int(total), list(off), list(res), list(sc)


def mlp_param_count(in_dim, width, n_hidden):
    return int(_lib.lib.mfn_mlp_param_count(int(in_dim), int(width), int(n_hidden)))


def _dev_guard(t):
    if not t.is_cuda:
        raise RuntimeError("mfnerf_b200 field ops need CUDA tensors (there is no CPU fallback)")
    return torch.cuda.device(t.device)


def grid_encode_fwd(x01, table_h, cfg):
    n = x01.shape[0]
    out = torch.empty(n, cfg.n_levels * cfg.n_features, dtype=torch.float16, device=x01.device)
    with _dev_guard(x01):
        call("mfn_grid_encode_fwd", ptr(x01), ptr(table_h), ctypes.byref(cfg), n, ptr(out), stream_ptr(x01.device))
    return out


def grid_encode_bwd(x01, dL_dout_h, cfg, dgrid_f32):
    with _dev_guard(x01):
        call("mfn_grid_encode_bwd", ptr(x01), ptr(dL_dout_h), ctypes.byref(cfg), x01.shape[0], ptr(dgrid_f32), stream_ptr(x01.device))


def grid_encode_bwd_input(x01, table_h, dL_dout_h, cfg):
    """gradient w.r.t. the positions: dL_dout (n, L*F) fp16 -> (n,3) f32, same scale as dL_dout"""
    dx = torch.empty(x01.shape[0], 3, dtype=torch.float32, device=x01.device)
    with _dev_guard(x01):
        call("mfn_grid_encode_bwd_input", ptr(x01), ptr(table_h), ptr(dL_dout_h), ctypes.byref(cfg), x01.shape[0], ptr(dx), stream_ptr(x01.device))
    return dx


def sh4_bwd(d01, dL_dout_h):
    dd = torch.empty(d01.shape[0], 3, dtype=torch.float32, device=d01.device)
    with _dev_guard(d01):
        call("mfn_sh4_bwd", ptr(d01), ptr(dL_dout_h), d01.shape[0], ptr(dd), stream_ptr(d01.device))
    return dd


def sh4_fwd(d01, out=None, out_offset=0):
    n = d01.shape[0]
    if out is None:
        out = torch.empty(n, 16, dtype=torch.float16, device=d01.device)
    with _dev_guard(d01):
        call("mfn_sh4_fwd", ptr(d01), n, ptr(out), out.shape[1], int(out_offset), stream_ptr(d01.device))
    return out


def mlp_fwd(x_h, w_h, in_dim, width, n_hidden, act, save_acts=True):
    n = x_h.shape[0]
    out = torch.empty(n, 16, dtype=torch.float16, device=x_h.device)
    acts = torch.empty(n_hidden, n, width, dtype=torch.float16, device=x_h.device) if save_acts else None
    with _dev_guard(x_h):
        call("mfn_mlp_fwd", ptr(x_h), ptr(w_h), in_dim, width, n_hidden, ACT[act], n, ptr(out), ptr(acts), stream_ptr(x_h.device))
    return out, acts


def mlp_bwd(dout_h, x_h, acts, out_h, w_h, in_dim, width, n_hidden, act, dW_f32, need_dx=True):
    n = x_h.shape[0]
    dx = torch.empty(n, in_dim, dtype=torch.float16, device=x_h.device) if need_dx else None
    with _dev_guard(x_h):
        call("mfn_mlp_bwd", ptr(dout_h), ptr(x_h), ptr(acts), ptr(out_h), ptr(w_h), in_dim, width, n_hidden, ACT[act], n, ptr(dx),
             ptr(dW_f32), stream_ptr(x_h.device))
    return dx


def geo_cfg(grid_cfg, width, n_hidden):
    """field config of a bare grid + network module whose inputs are already in [0,1] (tcnn.NetworkWithInputEncoding)"""
    cfg = FieldCfg()
    cfg.grid = grid_cfg
    cfg.sigma_width, cfg.sigma_hidden = int(width), int(n_hidden)
    cfg.rgb_width, cfg.rgb_hidden, cfg.rgb_act = 64, 1, 0
    for k in range(3):
        cfg.xyz_min[k] = 0.0; cfg.xyz_max[k] = 1.0
    return cfg


def geo_fused(cfg):
    """True when mfn_geo_fwd runs this shape on the fused tcgen05 kernel"""
    return int(_lib.lib.mfn_field_is_fused(ctypes.byref(cfg))) == 1


def geo_fwd(x01, params_h, cfg):
    """x01 (n,3) f32, params_h = fp16 [3072 network | table] -> (n,16) fp16 raw outputs, one fused kernel, nothing saved"""
    n = x01.shape[0]
    out = torch.empty(n, 16, dtype=torch.float16, device=x01.device)
    with _dev_guard(x01):
        call("mfn_geo_fwd", ctypes.byref(cfg), ptr(params_h), ptr(x01), n, None, ptr(out), stream_ptr(x01.device))
    return out
