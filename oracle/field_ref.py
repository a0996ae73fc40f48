"""TEST INFRASTRUCTURE ONLY -- plain-torch restatement ("transliteration") of the tiny-cuda-nn pieces MF-NeRF uses:
multiresolution hash-grid encoding, degree-4 spherical harmonics and the bias-free fully fused MLPs
(call sites: /root/reference/models/networks.py:36-57, 60-67, 69-79; calls at :106, :146, :147).

PARITY UNPINNED: tiny-cuda-nn (an unpublished MF-NeRF fork of NVlabs/tiny-cuda-nn, no version pin, README.md:41)
is NOT in the reference tree and cannot be installed here, and the reference has no tests or golden vectors for this
boundary.  This file restates upstream tcnn's published algorithm (grid.h / spherical_harmonics.h /
fully_fused_mlp.cu semantics as summarised in SURVEY.md section 8c); it is the checker for mf-nerf_b200's encoder/MLP kernels
and the "torch transliteration" CPU baseline of BASELINE.md.  Runs on CPU or GPU; gradients come from torch autograd.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)


def grid_layout(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale, grid_type="Hash", n_tables=1):
    """per-level (offset, entries, resolution, scale).  tcnn: scale = exp2(l*log2(b))*N_min - 1 (b held as float),
    res = ceil(scale)+1, entries = min(round_up(res^3, 8), 2^T).
    grid_type "MixedFeature" (spec defined in include/mfnerf_b200.h, parity unpinned): level l lives in table k = l*K/L of 2^T
    entries, always hashed, vertices hashed by their coordinates in the table's canonical (finest) level."""
    log2b = math.log2(float(np.float32(per_level_scale)))
    off, levels = 0, []
    for l in range(n_levels):
        scale = float(np.float32(2.0 ** (l * log2b) * base_resolution - 1.0))
        res = int(math.ceil(scale)) + 1
        cells = res ** 3
        entries = min((min(cells, 0x7fffffff) + 7) // 8 * 8, 1 << log2_hashmap_size)
        levels.append(dict(offset=off, entries=entries, res=res, scale=scale, hashed=cells > entries))
        off += entries
    if grid_type == "MixedFeature":
        T = 1 << log2_hashmap_size
        for l, lv in enumerate(levels):
            k = l * n_tables // n_levels
            lc = max(j for j in range(n_levels) if j * n_tables // n_levels == k)
            lv.update(offset=k * T, entries=T, hashed=True, canon=float(np.float32(levels[lc]["scale"]) / np.float32(lv["scale"])))
        off = n_tables * T
    elif grid_type != "Hash":
        raise NotImplementedError(grid_type)
    return levels, off


def _canon(v, ratio):
    """MixedFeature index transformation: level vertex -> nearest vertex of the table's canonical grid, in fp32 like the kernel:
    floor(fl(fl(v - 0.5) * ratio) + 1)"""
    r = torch.tensor(ratio, dtype=torch.float32, device=v.device)
    return torch.floor((v.to(torch.float32) - 0.5) * r + 1.0).to(torch.int64)


def _corner_indices(gx, gy, gz, lv):
    """(N,) int64 coords -> (N,) int64 table index within the level (tcnn grid_index)"""
    if "canon" in lv:
        gx, gy, gz = _canon(gx, lv["canon"]), _canon(gy, lv["canon"]), _canon(gz, lv["canon"])
    if lv["hashed"]:
        m = 0xFFFFFFFF
        idx = ((gx * PRIMES[0]) & m) ^ ((gy * PRIMES[1]) & m) ^ ((gz * PRIMES[2]) & m)
    else:
        idx = (gx + gy * lv["res"] + gz * lv["res"] * lv["res"]) & 0xFFFFFFFF
    return idx % lv["entries"]


def grid_encode(x01, table, levels, n_features):
    """x01 (N,3) float32 in [0,1]; table (entries*F,) float (fp16-representable values) -> (N, L*F) float32 (un-rounded).
    Differentiable w.r.t. `table`."""
    N = x01.shape[0]
    tab = table.view(-1, n_features)
    outs = []
    for lv in levels:
        pos = torch.addcmul(torch.full_like(x01, 0.5), x01, torch.tensor(lv["scale"], dtype=x01.dtype, device=x01.device))
        fl = torch.floor(pos)
        w = pos - fl
        g = fl.to(torch.int64)
        acc = torch.zeros(N, n_features, dtype=torch.float32, device=x01.device)
        for c in range(8):
            dx, dy, dz = c & 1, (c >> 1) & 1, c >> 2
            wc = (w[:, 0] if dx else 1 - w[:, 0]) * (w[:, 1] if dy else 1 - w[:, 1]) * (w[:, 2] if dz else 1 - w[:, 2])
            idx = _corner_indices(g[:, 0] + dx, g[:, 1] + dy, g[:, 2] + dz, lv) + lv["offset"]
            acc = acc + wc[:, None] * tab[idx].float()
        outs.append(acc)
    return torch.cat(outs, 1)


def sh4(d01):
    """degree-4 real spherical harmonics of 2*d01-1 -> (N,16) float32 (tcnn SphericalHarmonics encoding)"""
    v = d01.float() * 2 - 1
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([
        torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], 1)


def _h(t):
    """round to fp16 and come back (identity gradient): the points where tcnn holds activations in half precision"""
    return t + (t.half().float() - t).detach()


def mlp_split(weights, in_dim, width, n_hidden, out_pad=16):
    shapes = [(width, in_dim)] + [(width, width)] * (n_hidden - 1) + [(out_pad, width)]
    mats, o = [], 0
    for r, c in shapes:
        mats.append(weights[o:o + r * c].view(r, c)); o += r * c
    return mats


def mlp(x, weights, in_dim, width, n_hidden, out_act="None", return_hidden=False):
    """x (N,in_dim) (fp16-representable), weights flat (fp16-representable) -> (N,16) float32 un-rounded output
    (tcnn FullyFusedMLP: no biases, ReLU, fp16 activations between layers; accumulation here is fp32)."""
    mats = mlp_split(weights.float(), in_dim, width, n_hidden)
    h = x.float()
    hidden = []
    for W in mats[:-1]:
        h = _h(torch.relu(h @ W.t()))
        hidden.append(h)
    o = h @ mats[-1].t()
    if out_act == "Sigmoid":
        o = torch.sigmoid(o)
    elif out_act == "Exponential":
        o = torch.exp(o)
    return (o, hidden) if return_hidden else o


class NGPRef(torch.nn.Module):
    """torch restatement of models/networks.py:12-155 (NGP.density / NGP.forward) on top of the functions above.
    Parameters are fp32 masters; forward uses their fp16-rounded values like tcnn does."""

    def __init__(self, scale, L=16, F=2, log2_T=19, N_min=16, N_max=2048, rgb_channels=64, rgb_layers=2, params=None, grid="Hash", n_tables=1):
        super().__init__()
        self.scale = scale
        self.L, self.F = L, F
        self.b = float(np.exp(np.log(N_max * scale / N_min) / (L - 1)))
        self.levels, self.entries = grid_layout(L, F, log2_T, N_min, self.b, grid, n_tables)
        self.rgb_channels, self.rgb_layers = rgb_channels, rgb_layers
        n_xyz = 64 * (L * F) + 16 * 64 + self.entries * F
        n_rgb = rgb_channels * 32 + (rgb_layers - 1) * rgb_channels ** 2 + 16 * rgb_channels
        if params is None:
            g = torch.Generator().manual_seed(1337)
            xyz = torch.empty(n_xyz).uniform_(-1e-4, 1e-4, generator=g)
            xyz[:64 * L * F + 1024].uniform_(-0.25, 0.25, generator=g)
            rgb = torch.empty(n_rgb).uniform_(-0.25, 0.25, generator=g)
        else:
            xyz, rgb = params
        self.xyz_params = torch.nn.Parameter(xyz.detach().clone().float())
        self.rgb_params = torch.nn.Parameter(rgb.detach().clone().float())

    def density_h(self, x):
        x01 = (x - (-self.scale)) / (self.scale - (-self.scale))
        n_mlp = 64 * self.L * self.F + 1024
        p = _h(self.xyz_params)
        feats = _h(grid_encode(x01, p[n_mlp:], self.levels, self.F))
        h = _h(mlp(feats, p[:n_mlp], self.L * self.F, 64, 1))
        return torch.exp(h[:, 0]), h

    def forward(self, x, d):
        sigmas, h = self.density_h(x)
        d = d / torch.norm(d, dim=1, keepdim=True)
        enc = _h(sh4((d + 1) / 2))
        rgbs = _h(mlp(torch.cat([enc, h], 1), _h(self.rgb_params), 32, self.rgb_channels, self.rgb_layers, "Sigmoid"))[:, :3]
        return sigmas, rgbs
