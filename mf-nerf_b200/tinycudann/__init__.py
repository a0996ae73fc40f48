"""Drop-in for the subset of tiny-cuda-nn's PyTorch API that MF-NeRF uses (models/networks.py:36-94):
`NetworkWithInputEncoding`, `Encoding`, `Network`, each with a flat fp32 `.params` nn.Parameter, fp16 outputs and
tcnn's loss-scale-128 backward convention.  The kernels are mfnerf_b200's own (hash-grid encoder, SH, fused MLPs).

tcnn itself is NOT part of the reference tree (un-vendored, unpinned fork), so the arithmetic here follows
upstream tcnn's published algorithm -- parity for this boundary is "unpinned".
Parameter layout of NetworkWithInputEncoding: [MLP weights | grid table]  (networks.py:58 relies on the
first 3072 entries being the 32->64->16 MLP).
"""
import math

import torch
from torch import nn

from mfnerf_b200 import field_ops as F

__all__ = ["NetworkWithInputEncoding", "Encoding", "Network"]

LOSS_SCALE = 128.0


def _xavier_uniform_(w, fan_out, fan_in, gen):
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return w.uniform_(-bound, bound, generator=gen)


def _init_mlp(params, in_dim, width, n_hidden, n_out_padded, gen):
    """tcnn FullyFusedMLP init: Xavier-uniform per weight matrix, matrices stored first -> last, row-major (out x in)"""
    o = 0
    shapes = [(width, in_dim)] + [(width, width)] * (n_hidden - 1) + [(n_out_padded, width)]
    for (r, c) in shapes:
        _xavier_uniform_(params[o:o + r * c], r, c, gen)
        o += r * c
    return o


class _ParamCache:
    """fp16 shadow of the fp32 master parameters, re-cast on EVERY forward (one cast kernel, ~11 us for the 11.4 M parameters of
    the Lego configuration).  It must not be keyed on `Parameter._version`: the reference trainer's optimiser is apex FusedAdam
    (train.py:23,136), which updates `p.data` through raw pointers and never bumps the version counter -- a version-keyed cache
    would keep serving the initial weights while the fp32 masters drift.  The autograd Functions get a fresh tensor (they save it
    for backward); the no-grad fast path (NGP.density in update_density_grid, the test-time render loop) reuses one buffer."""

    def __init__(self):
        self._buf = None

    def get(self, p, fresh=True):
        if fresh:
            return p.detach().to(torch.float16)
        if self._buf is None or self._buf.shape != p.shape or self._buf.device != p.device:
            self._buf = torch.empty_like(p, dtype=torch.float16)
        self._buf.copy_(p.detach())
        return self._buf


def _parse_network(cfg):
    if cfg.get("otype", "FullyFusedMLP") not in ("FullyFusedMLP", "CutlassMLP"):
        raise NotImplementedError(f"network otype {cfg.get('otype')!r}")
    if cfg.get("activation", "ReLU") != "ReLU":
        raise NotImplementedError("only ReLU hidden activations are implemented")
    act = cfg.get("output_activation", "None")
    if act not in F.ACT:
        raise NotImplementedError(f"output activation {act!r}")
    return int(cfg["n_neurons"]), int(cfg["n_hidden_layers"]), act


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, mod):
        xh = mod._prep_input(x)
        wh = mod._cache.get(params)
        out, acts = F.mlp_fwd(xh, wh, mod.in_padded, mod.width, mod.n_hidden, mod.out_act, save_acts=True)
        ctx.mod = mod
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(xh, acts, out, wh)
        return out[:, :mod.n_output_dims]

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        xh, acts, out, wh = ctx.saved_tensors
        d16 = torch.zeros(xh.shape[0], 16, dtype=torch.float16, device=xh.device)
        d16[:, :mod.n_output_dims] = (dout.float() * mod.loss_scale).to(torch.float16)
        dW = torch.zeros(wh.shape[0], dtype=torch.float32, device=xh.device)
        dx = F.mlp_bwd(d16, xh, acts, out, wh, mod.in_padded, mod.width, mod.n_hidden, mod.out_act, dW, need_dx=ctx.needs_input_grad[0])
        dW /= mod.loss_scale
        if dx is not None:
            dx = (dx[:, :mod.n_input_dims].float() / mod.loss_scale).to(ctx.x_dtype)
        return dx, dW, None


class Network(nn.Module):
    """tcnn.Network(n_input_dims, n_output_dims, network_config) -- networks.py:69-79, 83-94"""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.width, self.n_hidden, self.out_act = _parse_network(network_config)
        self.in_padded = (self.n_input_dims + 15) // 16 * 16
        if self.n_output_dims > 16:
            raise NotImplementedError("n_output_dims > 16")
        self.loss_scale = LOSS_SCALE
        n = F.mlp_param_count(self.in_padded, self.width, self.n_hidden)
        gen = torch.Generator().manual_seed(seed)
        p = torch.empty(n, dtype=torch.float32)
        _init_mlp(p, self.in_padded, self.width, self.n_hidden, 16, gen)
        self.params = nn.Parameter(p)
        self._cache = _ParamCache()

    def _prep_input(self, x):
        x = x.to(torch.float16)
        if self.in_padded != self.n_input_dims:  # tcnn pads the (identity-encoded) input with ones
            x = torch.cat([x, torch.ones(x.shape[0], self.in_padded - self.n_input_dims, dtype=x.dtype, device=x.device)], 1)
        return x.contiguous()

    def forward(self, x):
        return _MlpFn.apply(x, self.params, self)


class _ShFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xf = x.float().contiguous()
        ctx.save_for_backward(xf)
        ctx.x_dtype = x.dtype
        return F.sh4_fwd(xf)

    @staticmethod
    def backward(ctx, dout):
        # gradient w.r.t. the view direction: the reference's --optimize_ext path (custom_functions.py:102-112, train.py:91,122,138)
        (xf,) = ctx.saved_tensors
        return F.sh4_bwd(xf, dout.to(torch.float16).contiguous()).to(ctx.x_dtype)


def _grid_input_grad(xf, table_h, dfeats_h, mod, x_dtype):
    """dL/dx01 from the (loss-scaled, fp16) gradient of the encoded features -- the --optimize_ext path"""
    return (F.grid_encode_bwd_input(xf, table_h, dfeats_h, mod.grid_cfg) / mod.loss_scale).to(x_dtype)


class _GridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, mod):
        xf = x.float().contiguous()
        ph = mod._cache.get(params)
        feats = F.grid_encode_fwd(xf, ph, mod.grid_cfg)
        ctx.mod = mod
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(xf, ph)
        return feats

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        xf, ph = ctx.saved_tensors
        d16 = (dout.float() * mod.loss_scale).to(torch.float16).contiguous()
        dgrid = None
        if ctx.needs_input_grad[1]:
            dgrid = torch.zeros(mod.params.shape[0], dtype=torch.float32, device=xf.device)
            F.grid_encode_bwd(xf, d16, mod.grid_cfg, dgrid)
            dgrid /= mod.loss_scale
        dx = _grid_input_grad(xf, ph, d16, mod, ctx.x_dtype) if ctx.needs_input_grad[0] else None
        return dx, dgrid, None


def _parse_grid(cfg):
    otype = cfg.get("otype", "HashGrid")
    gtype = cfg.get("type", otype.replace("Grid", ""))
    if cfg.get("interpolation", "Linear") != "Linear":
        raise NotImplementedError("only Linear interpolation is implemented")
    return F.make_grid_cfg(cfg.get("n_levels", 16), cfg.get("n_features_per_level", 2), cfg.get("log2_hashmap_size", 19),
                           cfg.get("base_resolution", 16), cfg.get("per_level_scale", 2.0), gtype, cfg.get("n_tables", 1))


class Encoding(nn.Module):
    """tcnn.Encoding(n_input_dims, encoding_config) -- networks.py:60-67 (SphericalHarmonics deg 4) or a grid"""

    def __init__(self, n_input_dims, encoding_config, seed=1337):
        super().__init__()
        self.n_input_dims = int(n_input_dims)
        self.otype = encoding_config["otype"]
        self.loss_scale = LOSS_SCALE
        self._cache = _ParamCache()
        if self.otype == "SphericalHarmonics":
            if int(encoding_config.get("degree", 4)) != 4 or self.n_input_dims != 3:
                raise NotImplementedError("SphericalHarmonics: only degree 4 on 3 inputs")
            self.n_output_dims = 16
            self.params = nn.Parameter(torch.zeros(0, dtype=torch.float32))
        elif self.otype.endswith("Grid"):
            self.grid_cfg = _parse_grid(encoding_config)
            total, *_ = F.grid_layout(self.grid_cfg)
            self.n_output_dims = self.grid_cfg.n_levels * self.grid_cfg.n_features
            gen = torch.Generator().manual_seed(seed)
            self.params = nn.Parameter(torch.empty(total * self.grid_cfg.n_features, dtype=torch.float32).uniform_(-1e-4, 1e-4, generator=gen))
        else:
            raise NotImplementedError(f"encoding otype {self.otype!r}")

    def forward(self, x):
        if self.otype == "SphericalHarmonics":
            return _ShFn.apply(x)
        return _GridFn.apply(x, self.params, self)


class _EncMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, mod):
        xf = x.float().contiguous()
        ph = mod._cache.get(params)
        w_mlp, table = ph[:mod.n_mlp_params], ph[mod.n_mlp_params:]
        feats = F.grid_encode_fwd(xf, table, mod.grid_cfg)
        out, acts = F.mlp_fwd(feats, w_mlp, mod.n_enc_out, mod.width, mod.n_hidden, mod.out_act, save_acts=True)
        ctx.mod = mod
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(xf, feats, acts, out, ph)
        return out[:, :mod.n_output_dims]

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        xf, feats, acts, out, ph = ctx.saved_tensors
        n = xf.shape[0]
        d16 = torch.zeros(n, 16, dtype=torch.float16, device=xf.device)
        d16[:, :mod.n_output_dims] = (dout.float() * mod.loss_scale).to(torch.float16)
        dparams = torch.zeros(ph.shape[0], dtype=torch.float32, device=xf.device)
        dfeats = F.mlp_bwd(d16, feats, acts, out, ph[:mod.n_mlp_params], mod.n_enc_out, mod.width, mod.n_hidden, mod.out_act,
                           dparams[:mod.n_mlp_params], need_dx=True)
        F.grid_encode_bwd(xf, dfeats, mod.grid_cfg, dparams[mod.n_mlp_params:])
        dparams /= mod.loss_scale
        dx = _grid_input_grad(xf, ph[mod.n_mlp_params:], dfeats, mod, ctx.x_dtype) if ctx.needs_input_grad[0] else None
        return dx, dparams, None


class NetworkWithInputEncoding(nn.Module):
    """tcnn.NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config) -- networks.py:36-57"""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        if int(n_input_dims) != 3:
            raise NotImplementedError("grid encodings take 3-D positions")
        self.n_input_dims, self.n_output_dims = 3, int(n_output_dims)
        self.grid_cfg = _parse_grid(encoding_config)
        self.width, self.n_hidden, self.out_act = _parse_network(network_config)
        self.n_enc_out = self.grid_cfg.n_levels * self.grid_cfg.n_features
        if self.n_enc_out not in (16, 32, 64):
            raise NotImplementedError("n_levels * n_features_per_level must be 16, 32 or 64")
        self.loss_scale = LOSS_SCALE
        self.n_mlp_params = F.mlp_param_count(self.n_enc_out, self.width, self.n_hidden)
        total, *_ = F.grid_layout(self.grid_cfg)
        gen = torch.Generator().manual_seed(seed)
        p = torch.empty(self.n_mlp_params + total * self.grid_cfg.n_features, dtype=torch.float32)
        _init_mlp(p, self.n_enc_out, self.width, self.n_hidden, 16, gen)
        p[self.n_mlp_params:].uniform_(-1e-4, 1e-4, generator=gen)
        self.params = nn.Parameter(p)
        self._cache = _ParamCache()
        self._geo_cfg = F.geo_cfg(self.grid_cfg, self.width, self.n_hidden)
        self._geo_fused = None

    def forward(self, x):
        if not (torch.is_grad_enabled() and (self.params.requires_grad or x.requires_grad)):
            # inference (NGP.density in update_density_grid, the whole test-time render): gather + both layers in ONE tcgen05
            # kernel, no activations written.  Raw outputs only -- an output activation would need the unfused path.
            if self._geo_fused is None:
                self._geo_fused = F.geo_fused(self._geo_cfg) and self.out_act in ("None", "none", None)
            if self._geo_fused and x.is_cuda:
                out = F.geo_fwd(x.detach().float().contiguous(), self._cache.get(self.params, fresh=False), self._geo_cfg)
                return out[:, :self.n_output_dims]
        return _EncMlpFn.apply(x, self.params, self)
