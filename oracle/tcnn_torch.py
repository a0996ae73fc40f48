"""BASELINE INFRASTRUCTURE ONLY -- a torch transliteration of the three tiny-cuda-nn module classes MF-NeRF constructs
(/root/reference/models/networks.py:36-79), importable AS `tinycudann` by the reference's unmodified python layer, so that
bench.py's `gpu_reference` leg can time "GPU-ref A+B" (BASELINE.md section 3): the reference's own compiled vren kernels
(oracle/_ref/vren_ref*.so) + this stand-in for the un-vendored tcnn fork, driven by the reference's rendering.py / losses.py /
custom_functions.py.  It is NOT tiny-cuda-nn (which cannot be installed here: not in the tree, no network) and is always
labelled "torch transliteration"; tcnn's fused kernels would be faster than this -- see DESIGN.md for how the ratio is read.

Semantics follow oracle/field_ref.py (same grid layout, hash, interpolation, SH, bias-free ReLU MLPs, fp16 activations); the
implementation is the fastest plain-torch form we know: per level one (N, 8) index computation and gather, fp16 cuBLAS GEMMs for
the MLPs, and a custom backward that re-derives the indices and scatters with index_add_ (float atomics) instead of autograd's
sort-based index_put.  PARITY UNPINNED like field_ref.py.  Only bench.py (gpu_reference leg) and tests/ may import it.
"""
import math

import numpy as np
import torch
from torch import nn

from . import field_ref as fr

_OFFS = None


def _corner_offsets(device):
    return torch.tensor([[c & 1, (c >> 1) & 1, c >> 2] for c in range(8)], dtype=torch.int64, device=device)   # (8,3)


def _level_indices_weights(x01, lv):
    """x01 (N,3) f32 -> idx (N,8) int64 into the flat table (entries), w (N,8) f32"""
    pos = torch.addcmul(torch.full_like(x01, 0.5), x01, torch.tensor(lv["scale"], dtype=x01.dtype, device=x01.device))
    fl = torch.floor(pos)
    w = pos - fl
    g = fl.to(torch.int64)
    c = g[:, None, :] + _corner_offsets(x01.device)[None]                      # (N,8,3)
    if "canon" in lv:
        c = torch.floor((c.to(torch.float32) - 0.5) * torch.tensor(lv["canon"], dtype=torch.float32, device=x01.device) + 1.0).to(torch.int64)
    if lv["hashed"]:
        m = 0xFFFFFFFF
        idx = ((c[..., 0] * fr.PRIMES[0]) & m) ^ ((c[..., 1] * fr.PRIMES[1]) & m) ^ ((c[..., 2] * fr.PRIMES[2]) & m)
    else:
        idx = (c[..., 0] + c[..., 1] * lv["res"] + c[..., 2] * (lv["res"] * lv["res"])) & 0xFFFFFFFF
    idx = idx % lv["entries"] + lv["offset"]
    offs = _corner_offsets(x01.device).to(torch.bool)                          # (8,3)
    w3 = torch.where(offs[None], w[:, None, :], 1.0 - w[:, None, :])           # (N,8,3)
    return idx, w3[..., 0] * w3[..., 1] * w3[..., 2]


class _GridEncode(torch.autograd.Function):
    """x01 (N,3), flat fp32 table (entries*F) -> (N, L*F) fp16; backward re-derives the indices (like tcnn) and index_add_s in fp32"""

    @staticmethod
    def forward(ctx, x01, table, levels, F):
        x01 = x01.float()
        tab = table.detach().to(torch.float16).view(-1, F)
        outs = []
        for lv in levels:
            idx, w = _level_indices_weights(x01, lv)
            outs.append((tab[idx].float() * w[..., None]).sum(1))
        ctx.levels, ctx.F, ctx.n_table = levels, F, table.shape[0]
        ctx.save_for_backward(x01)
        return torch.cat(outs, 1).to(torch.float16)

    @staticmethod
    def backward(ctx, dout):
        (x01,) = ctx.saved_tensors
        F = ctx.F
        dtab = torch.zeros(ctx.n_table // F, F, dtype=torch.float32, device=dout.device)
        d = dout.float()
        for l, lv in enumerate(ctx.levels):
            idx, w = _level_indices_weights(x01, lv)
            dtab.index_add_(0, idx.reshape(-1), (w[..., None] * d[:, None, l * F:(l + 1) * F]).reshape(-1, F))
        return None, dtab.view(-1), None, None


def _mlp_half(x, weights, in_dim, width, n_hidden, out_act):
    mats = fr.mlp_split(weights.to(torch.float16), in_dim, width, n_hidden)
    h = x.to(torch.float16)
    for W in mats[:-1]:
        h = torch.relu(h @ W.t())
    o = h @ mats[-1].t()
    if out_act == "Sigmoid":
        o = torch.sigmoid(o)
    elif out_act == "Exponential":
        o = torch.exp(o)
    return o


def _init_mlp(p, in_dim, width, n_hidden, gen):
    o = 0
    for r, c in [(width, in_dim)] + [(width, width)] * (n_hidden - 1) + [(16, width)]:
        b = math.sqrt(6.0 / (r + c))
        p[o:o + r * c].uniform_(-b, b, generator=gen); o += r * c
    return o


def _grid_levels(cfg):
    gtype = cfg.get("type", cfg.get("otype", "HashGrid").replace("Grid", ""))
    return fr.grid_layout(cfg.get("n_levels", 16), cfg.get("n_features_per_level", 2), cfg.get("log2_hashmap_size", 19), cfg.get("base_resolution", 16),
                          cfg.get("per_level_scale", 2.0), gtype, cfg.get("n_tables", 1))


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        self.n_output_dims = int(n_output_dims)
        self.F = int(encoding_config.get("n_features_per_level", 2))
        self.levels, entries = _grid_levels(encoding_config)
        self.n_enc = len(self.levels) * self.F
        self.width, self.n_hidden = int(network_config["n_neurons"]), int(network_config["n_hidden_layers"])
        self.out_act = network_config.get("output_activation", "None")
        self.n_mlp = self.width * self.n_enc + (self.n_hidden - 1) * self.width ** 2 + 16 * self.width
        g = torch.Generator().manual_seed(seed)
        p = torch.empty(self.n_mlp + entries * self.F)
        _init_mlp(p, self.n_enc, self.width, self.n_hidden, g)
        p[self.n_mlp:].uniform_(-1e-4, 1e-4, generator=g)
        self.params = nn.Parameter(p)

    def forward(self, x):
        feats = _GridEncode.apply(x, self.params[self.n_mlp:], self.levels, self.F)
        return _mlp_half(feats, self.params[:self.n_mlp], self.n_enc, self.width, self.n_hidden, self.out_act)[:, :self.n_output_dims]


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337):
        super().__init__()
        if encoding_config["otype"] != "SphericalHarmonics":
            raise NotImplementedError(encoding_config["otype"])
        self.n_output_dims = 16
        self.params = nn.Parameter(torch.zeros(0))

    def forward(self, x):
        return fr.sh4(x).to(torch.float16)


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.in_padded = (self.n_input_dims + 15) // 16 * 16
        self.width, self.n_hidden = int(network_config["n_neurons"]), int(network_config["n_hidden_layers"])
        self.out_act = network_config.get("output_activation", "None")
        g = torch.Generator().manual_seed(seed)
        p = torch.empty(self.width * self.in_padded + (self.n_hidden - 1) * self.width ** 2 + 16 * self.width)
        _init_mlp(p, self.in_padded, self.width, self.n_hidden, g)
        self.params = nn.Parameter(p)

    def forward(self, x):
        if self.in_padded != self.n_input_dims:
            x = torch.cat([x, torch.ones(x.shape[0], self.in_padded - self.n_input_dims, dtype=x.dtype, device=x.device)], 1)
        return _mlp_half(x, self.params, self.in_padded, self.width, self.n_hidden, self.out_act)[:, :self.n_output_dims]
