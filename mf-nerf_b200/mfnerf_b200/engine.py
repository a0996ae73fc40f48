"""Sync-free training / rendering engine: the reference's `training_step` (train.py:164-190) and
`render(test_time=True)` (rendering.py:46-118) driven directly through the C ABI, without autograd and without
host synchronisation inside a step.

Step = [occupancy update every 16 steps] -> AABB + near clamp + jitter -> march -> field fwd -> composite fw -> loss (+ distortion)
       -> composite bw -> field bwd + hash-grid scatter -> [gradient exchange when world_size > 1] -> fused Adam.
Sample arrays live in fixed-capacity buffers and every kernel reads the live sample count from device memory, so the step is
three CUDA graphs (marching front | field front | field back) replayed on three streams as a pipeline: the marching front of step
t+1 overlaps the field backward, scatter and optimiser of step t (`_step_dp`).  `pipelined=False` runs the same kernels on one
stream (tests compare the two).  Rendering is the device-side wavefront of csrc/render.cu.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, field_ops
from . import dist as mdist
from ._lib import call, ptr, stream_ptr

MAX_SAMPLES = 1024          # rendering.py:7
NEAR_DISTANCE = 0.01        # rendering.py:8
G = 128                     # networks.py:27


FieldCfg = field_ops.FieldCfg


def make_field_cfg(scale, L=16, F=2, log2_T=19, N_min=16, N_max=2048, rgb_channels=64, rgb_layers=2, rgb_act="Sigmoid", grid="Hash",
                   n_tables=1):
    b = float(np.exp(np.log(N_max * scale / N_min) / (L - 1)))       # networks.py:33
    cfg = FieldCfg()
    cfg.grid = field_ops.make_grid_cfg(L, F, log2_T, N_min, b, grid, n_tables)
    cfg.sigma_width, cfg.sigma_hidden = 64, 1
    cfg.rgb_width, cfg.rgb_hidden, cfg.rgb_act = int(rgb_channels), int(rgb_layers), field_ops.ACT[rgb_act]
    for k in range(3):
        cfg.xyz_min[k] = -scale; cfg.xyz_max[k] = scale
    return cfg


def _pad(n, m=8):
    return (n + m - 1) // m * m


class NGPEngine:
    def __init__(self, scale=0.5, L=16, F=2, log2_T=19, N_min=16, N_max=2048, rgb_channels=64, rgb_layers=2, n_rays=8192,
                 device="cuda", lr=1e-2, loss_scale=1024.0, distortion_w=0.0, lambda_opacity=1e-3, sample_capacity=None, seed=1337,
                 exp_step_factor=None, T_threshold=1e-4, world_size=1, process_group=None, force_dp_path=False, pipelined=True,
                 grid="Hash", n_tables=1):
        self.dev = torch.device(device)
        self.scale, self.n_rays = float(scale), int(n_rays)
        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)                     # networks.py:26
        self.esf = (1.0 / 256 if scale > 0.5 else 0.0) if exp_step_factor is None else float(exp_step_factor)   # train.py:100-101
        self.bg = (ctypes.c_float * 3)(*([1.0, 1.0, 1.0] if self.esf == 0 else [0.0, 0.0, 0.0]))           # rendering.py:153-161
        self.cfg = make_field_cfg(scale, L, F, log2_T, N_min, N_max, rgb_channels, rgb_layers, grid=grid, n_tables=n_tables)     # opt.py:71-85 --grid / --N_tables
        self.lr, self.loss_scale, self.T_thr = lr, float(loss_scale), float(T_threshold)
        self.distortion_w, self.lambda_opacity = float(distortion_w), float(lambda_opacity)
        self.world_size, self.pg = world_size, process_group
        # pipelined: the optimiser (and, data-parallel, the gradient exchange) of step t runs on a side stream while step t+1's
        # parameter-independent front (batch copy, AABB, marching) is already under way -- same arithmetic, same order per datum.
        # force_dp_path: run the collectives of the data-parallel path on a 1-rank group (tests)
        self.collectives = world_size > 1 or bool(force_dp_path)
        self.dp = self.collectives or bool(pipelined)
        self._deep = False     # set in __init__ once the field config exists: next step's march may start before this step's backward
        entries, *_ = field_ops.grid_layout(self.cfg.grid)
        self.n_mlp1 = 64 * 32 + 16 * 64
        self.n_xyz = self.n_mlp1 + entries * F
        self.n_rgb = field_ops.mlp_param_count(32, rgb_channels, rgb_layers)
        self.off_rgb = _pad(self.n_xyz)
        self.n_params = _pad(self.off_rgb + _pad(self.n_rgb), 8 * max(1, int(world_size)))
        d = self.dev
        g = torch.Generator().manual_seed(seed)
        p = torch.zeros(self.n_params)
        from tinycudann import _init_mlp
        _init_mlp(p[:self.n_mlp1], 32, 64, 1, 16, g)
        p[self.n_mlp1:self.n_xyz].uniform_(-1e-4, 1e-4, generator=g)
        _init_mlp(p[self.off_rgb:self.off_rgb + self.n_rgb], 32, rgb_channels, rgb_layers, 16, g)
        self.params = p.to(d)
        # Data-parallel: gradients, fp16 shadow parameters and the overflow flag live in ONE symmetric-memory allocation (peer-mapped on
        # every rank; its NVSwitch multicast address is used on request only) so that the gradient exchange + optimiser is one kernel over NVLink
        # (csrc/dp_exchange.cu).  MFN_DP_EXCHANGE=nccl keeps the three NCCL collectives (A/B measurements, fabrics without P2P).
        self._symm = None
        if world_size > 1 and os.environ.get("MFN_DP_EXCHANGE", "fused") == "fused":
            self._symm = mdist.SymmetricBuffers(self.n_params, d, process_group)
        if self._symm is not None:
            self.params_h, self.grads, self.overflow = self._symm.shadow, self._symm.grads, self._symm.flag
            self.params_h.copy_(self.params); self.grads.zero_(); self.overflow.zero_()
        else:
            self.params_h = self.params.half()
            self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.step_count = 0
        # occupancy grid (train.py:78-81, networks.py:27-29)
        self.density_grid = torch.zeros(self.cascades, G ** 3, device=d)
        self.density_bitfield = torch.zeros(self.cascades * G ** 3 // 8, dtype=torch.uint8, device=d)
        from .synthetic import morton_order_coords
        self.cell_coords = torch.from_numpy(morton_order_coords(G)).to(d)      # row m = coords of morton index m
        # per-ray buffers
        R = self.n_rays
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=d)
        self.rays = f(3, R, 3)                                   # packed [rays_o | rays_d | target]: one copy per step
        self.rays_o, self.rays_d, self.target = self.rays[0], self.rays[1], self.rays[2]
        self.hit_cnt = torch.empty(R, dtype=torch.int32, device=d)
        self.hits_t = f(R, 1, 2); self.hits_idx = torch.empty(R, 1, dtype=torch.int64, device=d)
        self.noise = f(R)
        self.rays_a = torch.empty(R, 3, dtype=torch.int64, device=d)
        self.counter = torch.zeros(2, dtype=torch.int32, device=d)
        self.n_field = torch.zeros(1, dtype=torch.int32, device=d)      # the field kernels' own copy of counter[0] (the next march overwrites counter)
        self.total_samples = torch.empty(R, dtype=torch.int64, device=d)
        self.opacity, self.depth, self.rgb, self.rgb_final = f(R), f(R), f(R, 3), f(R, 3)
        self.dL_dopacity, self.dL_ddepth, self.dL_drgb = f(R), torch.zeros(R, device=d), f(R, 3)
        self.dist_loss, self.dL_ddist = f(R), f(R)
        self.loss_terms = torch.zeros(3, device=d)
        if self._symm is None:
            self.overflow = torch.zeros(1, dtype=torch.int32, device=d)
        self._skip = torch.zeros(1, dtype=torch.int32, device=d)  # fused exchange: the OR of all ranks' overflow flags (what mfn_amp_update reads)
        self._adam_hyper = torch.zeros(4, device=d)              # {lr, 1 - beta1^t, 1 - beta2^t}: the optimiser's per-step scalars when it runs inside a graph
        # AMP state on the device (GradScaler semantics: the reference trains under Lightning precision=16): {loss scale, growth tracker,
        # skipped steps, applied steps}.  The backward kernel scales by [0], Adam unscales by it and takes its bias corrections from
        # [3] (a skipped step does not advance t), mfn_amp_update halves the scale on overflow and doubles it after 2000 good steps.
        self._amp = torch.zeros(8, device=d)
        call("mfn_amp_init", ptr(self._amp), float(loss_scale), 0, 0.9, 0.999, stream_ptr(d))
        self._amp_rule = (0.5, 2.0, 2000, 1.0, 65536.0)           # backoff, growth, growth interval, min, max (torch.amp.GradScaler defaults)
        self._lr_on_device = None
        self.center = torch.zeros(1, 3, device=d); self.half_size = torch.full((1, 3), self.scale, device=d)
        # per-sample buffers at fixed capacity
        self.cap = int(sample_capacity) if sample_capacity else R * MAX_SAMPLES
        S = self.cap
        self.xyzs, self.dirs, self.deltas, self.ts = f(S, 3), f(S, 3), f(S), f(S)
        self.sigmas, self.rgbs, self.ws = f(S), f(S, 3), f(S)
        self.dL_dsigmas, self.dL_drgbs, self.dL_dws = f(S), f(S, 3), torch.zeros(S, device=d)
        self.ws_incl, self.wts_incl = (f(S), f(S)) if self.distortion_w > 0 else (None, None)
        self.march_ws = torch.empty(_lib.lib.mfn_march_train_workspace_bytes(R, MAX_SAMPLES), dtype=torch.uint8, device=d)
        self.march_ws[:256].zero_()                              # header: ray queue, running sample total (@8), call counter (@16)
        self._march_hdr = self.march_ws[:256].view(torch.int64)  # [1] = samples marched so far, [2] = marcher calls
        self._box = ((ctypes.c_float * 3)(0.0, 0.0, 0.0), (ctypes.c_float * 3)(self.scale, self.scale, self.scale))
        self.field_ws = torch.empty(_lib.lib.mfn_field_workspace_bytes(ctypes.byref(self.cfg), S, 1), dtype=torch.uint8, device=d)
        # fused field kernels: the backward pass reads only its workspace, so step t+1's marching front may overwrite the sample arrays
        # as soon as step t's compositor backward is done and overlap step t's field backward + scatter
        self._fused = _lib.lib.mfn_field_is_fused(ctypes.byref(self.cfg)) == 1
        self._deep = self.dp and self._fused
        # fused shapes without distortion loss: compositing forward + loss + compositing backward are ONE kernel, the field forward
        # leaves its sample count in the workspace for the backward pass, and no copy / memset node is left in the step
        self._fast_front = self._fused and self.distortion_w == 0
        self._comp_scratch = torch.zeros(4, device=d)
        self._n_back = ctypes.c_void_p(_lib.lib.mfn_field_count_ptr(ctypes.byref(self.cfg), ptr(self.field_ws), S)) if self._fused else None
        self._graph = None
        self._graph_adam = None
        # one GPU: the next step's marching front waits for the scatter kernel and overlaps the optimiser only (memory-bound, it leaves
        # the issue slots free) instead of competing with the backward / scatter kernels
        self._march_after_scatter = os.environ.get("MFN_MARCH_OVERLAP", "back") == "adam"
        self._cells_ws = None
        self.graph_replays = 0
        self.launches_per_forward_backward = 0
        self.fixed_noise = None     # tests: jitter noise supplied by the caller instead of drawn per step

    # ------------------------------------------------------------------------------------------------ views
    def samples_marched(self):
        """running total of samples emitted by the training marcher (device int64 scalar; the reference's train/rm_s numerator)"""
        return self._march_hdr[1]

    def reset_samples_marched(self):
        self._march_hdr[1].zero_()

    @property
    def xyz_params_h(self):
        return self.params_h[:self.n_xyz]

    @property
    def rgb_params_h(self):
        return self.params_h[self.off_rgb:self.off_rgb + self.n_rgb]

    def state_dict(self):
        """reference checkpoint keys (utils.py / SURVEY section 5): flat fp32 tcnn parameter vectors + occupancy bitfield"""
        p = self.gather_master_params()
        return {"xyz_encoder.params": p[:self.n_xyz].clone(), "rgb_net.params": p[self.off_rgb:self.off_rgb + self.n_rgb].clone(),
                "dir_encoder.params": torch.zeros(0, device=self.dev), "density_bitfield": self.density_bitfield.clone(),
                "density_grid": self.density_grid.clone()}

    def load_state_dict(self, sd):
        self._wait_comm()
        self.params[:self.n_xyz].copy_(sd["xyz_encoder.params"]); self.params[self.off_rgb:self.off_rgb + self.n_rgb].copy_(sd["rgb_net.params"])
        self.params_h.copy_(self.params)
        if "density_bitfield" in sd:
            self.density_bitfield.copy_(sd["density_bitfield"])
        if "density_grid" in sd:
            self.density_grid.copy_(sd["density_grid"])

    def save_checkpoint(self, path, slim=True, model_name="model"):
        """writes what the reference's trainer leaves behind (a Lightning checkpoint: {'state_dict': {'model.<key>': tensor}}); slim=True
        drops `density_grid` like utils.slim_ckpt (utils.py:30-41).  utils.load_ckpt(model, path) of the reference reads it back."""
        sd = {f"{model_name}.{k}": v.cpu() for k, v in self.state_dict().items() if not (slim and k == "density_grid")}
        torch.save({"state_dict": sd}, path)

    def load_checkpoint(self, ckpt, model_name="model"):
        """a reference checkpoint -- path or dict, Lightning ('state_dict' -> 'model.xyz_encoder.params' ...) or already slimmed -- with
        the key handling of utils.extract_model_state_dict (utils.py:4-18)"""
        if not isinstance(ckpt, dict):
            ckpt = torch.load(ckpt, map_location="cpu")
        if "state_dict" in ckpt:
            ckpt = ckpt["state_dict"]
        sd = {k[len(model_name) + 1:]: v for k, v in ckpt.items() if k.startswith(model_name)}
        for k in ("xyz_encoder.params", "rgb_net.params"):
            if k not in sd:
                raise KeyError(f"checkpoint has no '{model_name}.{k}'")
        if sd["xyz_encoder.params"].numel() != self.n_xyz or sd["rgb_net.params"].numel() != self.n_rgb:
            raise ValueError(f"checkpoint is for another configuration: xyz_encoder.params {sd['xyz_encoder.params'].numel()} (engine {self.n_xyz}), "
                             f"rgb_net.params {sd['rgb_net.params'].numel()} (engine {self.n_rgb})")
        self.load_state_dict({k: v.to(self.dev) for k, v in sd.items() if torch.is_tensor(v)})

    # ------------------------------------------------------------------------------------------------ field queries
    @torch.no_grad()
    def field(self, xyzs, dirs):
        """NGP.forward (networks.py:134-155) without autograd: (n,3),(n,3) -> sigmas (n) f32, rgbs (n,3) f32"""
        self._wait_comm()
        n = xyzs.shape[0]
        cfg = ctypes.byref(self.cfg)
        need = _lib.lib.mfn_field_workspace_bytes(cfg, n, 0)
        if self._cells_ws is None or self._cells_ws.numel() < need:
            self._cells_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        sig = torch.empty(n, device=self.dev); col = torch.empty(n, 3, device=self.dev)
        call("mfn_field_fwd", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(xyzs.contiguous()), ptr(dirs.contiguous()), n, None,
             ptr(sig), ptr(col), ptr(self._cells_ws), self._cells_ws.numel(), stream_ptr(self.dev))
        return sig, col

    # ------------------------------------------------------------------------------------------------ occupancy grid
    @torch.no_grad()
    def density(self, xyz):
        """NGP.density (networks.py:96-110) for (n,3) world positions -> sigmas (n)"""
        self._wait_comm()
        n = xyz.shape[0]
        need = _lib.lib.mfn_field_workspace_bytes(ctypes.byref(self.cfg), n, 0)
        ws = self.field_ws
        if need > ws.numel():
            if self._cells_ws is None or self._cells_ws.numel() < need:
                self._cells_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
            ws = self._cells_ws
        out = torch.empty(n, device=self.dev)
        call("mfn_density_fwd", ctypes.byref(self.cfg), ptr(self.xyz_params_h), ptr(xyz), n, None, ptr(out), ptr(ws), ws.numel(), stream_ptr(self.dev))
        return out

    @torch.no_grad()
    def update_density_grid(self, density_threshold=0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=False, decay=0.95):
        """networks.py:242-271 without its host syncs (nonzero / .item()) and as a handful of fused kernels (csrc/density_grid.cu):
        cells -> jittered positions -> density query -> decay/max update -> mean of the positive cells -> packbits with the threshold
        min(mean, density_threshold) read on the device.  Occupied cells are drawn with a cumsum + searchsorted instead of nonzero.
        Two halves: _dg_prepare (which cells, where inside them: needs the occupancy grid only) and _dg_apply (density query with the
        CURRENT parameters, grid update, bitfield) -- the pipelined step runs the first half while the previous step's backward pass
        and optimiser are still in flight."""
        prep = self._dg_prepare(density_threshold, warmup)
        self._wait_comm()
        self._dg_apply(prep, density_threshold, decay)

    @torch.no_grad()
    def _dg_prepare(self, density_threshold=0.01 * MAX_SAMPLES / 3 ** 0.5, warmup=False):
        d, st = self.dev, stream_ptr(self.dev)
        G3 = G ** 3
        M = G3 // 4
        if not hasattr(self, "_dg_xyz"):
            self._dg_xyz = [torch.empty(G3, 3, device=d) for _ in range(self.cascades)]     # positions of the queried cells (all cells: G^3, sampled: 2M)
            self._dg_idx = [torch.empty(2 * M, dtype=torch.int32, device=d) for _ in range(self.cascades)]
            self._dg_sig = torch.empty(G3, device=d)
            self._dg_scratch = torch.zeros(4, device=d)
            self._dg_mean = torch.zeros(1, device=d)
            self._dg_calls = 0
        plan = []
        for c in range(self.cascades):
            self._dg_calls += 1
            seed = (self._dg_calls * 0x9E3779B97F4A7C15 + 1337) & 0xFFFFFFFFFFFFFFFF     # same sequence on every rank: replicas stay identical
            grid_c, xyz, idx = self.density_grid[c], self._dg_xyz[c], self._dg_idx[c]
            if warmup:                                    # get_all_cells (networks.py:157-168)
                n, idx_ptr = G3, None
                call("mfn_grid_cell_positions", None, 0, n, c, self.scale, G, seed, None, ptr(xyz), st)
            else:                                         # sample_uniform_and_occupied_cells (networks.py:170-197)
                n, idx_ptr = 2 * M, ptr(idx)
                call("mfn_grid_cell_positions", None, 1, M, c, self.scale, G, seed, ptr(idx), ptr(xyz), st)
                cs = torch.cumsum(grid_c > density_threshold, 0, dtype=torch.int32)
                call("mfn_grid_draw_occupied", ptr(cs), G3, M, seed ^ 0x3333333333333333, ptr(idx[M:]), st)
                # morton-sorted cells are spatially coherent: the density query's hash-grid gathers run ~3x faster than on the raw draw
                # (the update rule does not depend on the order of the cells)
                idx.copy_(torch.sort(idx)[0])
                call("mfn_grid_cell_positions", ptr(idx), 0, 2 * M, c, self.scale, G, seed ^ 0x5555555555555555, None, ptr(xyz), st)
            plan.append((c, n, idx_ptr))
        return plan

    @torch.no_grad()
    def _dg_apply(self, plan, density_threshold=0.01 * MAX_SAMPLES / 3 ** 0.5, decay=0.95):
        d, st = self.dev, stream_ptr(self.dev)
        G3 = G ** 3
        for c, n, idx_ptr in plan:
            grid_c = self.density_grid[c]
            sig = self._dg_sig[:n]
            need = _lib.lib.mfn_field_workspace_bytes(ctypes.byref(self.cfg), n, 0)
            if getattr(self, "_dg_ws", None) is None or self._dg_ws.numel() < need:      # never the training workspace: the previous
                self._dg_ws = torch.empty(need, dtype=torch.uint8, device=d)               # step's backward may still be reading it
            ws = self._dg_ws
            call("mfn_density_fwd", ctypes.byref(self.cfg), ptr(self.xyz_params_h), ptr(self._dg_xyz[c]), n, None, ptr(sig), ptr(ws), ws.numel(), st)
            call("mfn_grid_update", ptr(grid_c), idx_ptr, ptr(sig), G3, n, float(decay), st)
        call("mfn_grid_mean_positive", ptr(self.density_grid), self.density_grid.numel(), ptr(self._dg_scratch), ptr(self._dg_mean), st)
        self._mean_density = self._dg_mean
        call("mfn_packbits_dev_thr", ptr(self.density_grid), self.density_bitfield.numel(), float(density_threshold), ptr(self._dg_mean),
             ptr(self.density_bitfield), st)

    # ------------------------------------------------------------------------------------------------ training step
    def _forward_backward(self):
        """everything between "rays are in self.rays_o/d/target" and "self.grads holds loss_scale * dL/dparams"; no host sync"""
        self._march()
        self._field_backward()

    def _march(self):
        """parameter-independent front of the step: AABB, near clamp, jitter noise, ray marching (needs rays + bitfield only)"""
        d, st, R, S = self.dev, stream_ptr(self.dev), self.n_rays, self.cap
        if self.fixed_noise is None:       # jitter drawn inside the front-end kernel (custom_functions.py:83)
            call("mfn_ray_setup", ptr(self.rays_o), ptr(self.rays_d), self._box[0], self._box[1], R, NEAR_DISTANCE, ptr(self.march_ws[16:]), ptr(self.hits_t),
                 ptr(self.noise), st)
        else:
            call("mfn_ray_setup", ptr(self.rays_o), ptr(self.rays_d), self._box[0], self._box[1], R, NEAR_DISTANCE, None, ptr(self.hits_t), None, st)
            self.noise.copy_(self.fixed_noise)
        call("mfn_raymarching_train", ptr(self.rays_o), ptr(self.rays_d), ptr(self.hits_t), ptr(self.density_bitfield), self.cascades, self.scale,
             self.esf, ptr(self.noise), G, MAX_SAMPLES, R, S, ptr(self.rays_a), ptr(self.xyzs), ptr(self.dirs), ptr(self.deltas), ptr(self.ts),
             ptr(self.counter), ptr(self.march_ws), self.march_ws.numel(), st)

    def _field_backward(self):
        """field forward, compositing, loss, and the whole backward pass into self.grads"""
        self._field_front()
        self._field_back()

    def _field_front(self):
        """field forward, compositing forward, loss, compositing backward: the last readers of the marcher's sample arrays"""
        d, st, R, S = self.dev, stream_ptr(self.dev), self.n_rays, self.cap
        cfg = ctypes.byref(self.cfg)
        if self._fast_front:
            call("mfn_field_fwd", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(self.xyzs), ptr(self.dirs), S, ptr(self.counter), ptr(self.sigmas),
                 ptr(self.rgbs), ptr(self.field_ws), self.field_ws.numel(), st)
            call("mfn_composite_loss_train", ptr(self.sigmas), ptr(self.rgbs), ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), ptr(self.target), self.T_thr, R, S,
                 self.bg, self.lambda_opacity, 1.0, ptr(self.total_samples), ptr(self.opacity), ptr(self.depth), ptr(self.rgb), ptr(self.ws), ptr(self.rgb_final),
                 ptr(self.dL_drgb), ptr(self.dL_dopacity), ptr(self.dL_dsigmas), ptr(self.dL_drgbs), ptr(self.loss_terms), ptr(self._comp_scratch),
                 ptr(self.overflow), st)
            return
        self.n_field.copy_(self.counter[:1])
        n_dev = ptr(self.n_field)
        call("mfn_field_fwd", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(self.xyzs), ptr(self.dirs), S, n_dev, ptr(self.sigmas),
             ptr(self.rgbs), ptr(self.field_ws), self.field_ws.numel(), st)
        call("mfn_composite_train_fw", ptr(self.sigmas), ptr(self.rgbs), ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), self.T_thr, R, S,
             ptr(self.total_samples), ptr(self.opacity), ptr(self.depth), ptr(self.rgb), ptr(self.ws), st)
        dist = None
        if self.distortion_w > 0:
            call("mfn_distortion_loss_fw", ptr(self.ws), ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), R, S, ptr(self.dist_loss), ptr(self.ws_incl),
                 ptr(self.wts_incl), st)
            dist = self.dist_loss
        self.loss_terms.zero_()
        call("mfn_nerf_loss_fwbw", ptr(self.rgb), ptr(self.opacity), ptr(self.target), ptr(dist), R, self.bg, self.lambda_opacity, self.distortion_w, 1.0,
             ptr(self.dL_drgb), ptr(self.dL_dopacity), ptr(self.dL_ddist) if dist is not None else None, ptr(self.rgb_final), ptr(self.loss_terms), st)
        if dist is not None:
            call("mfn_distortion_loss_bw", ptr(self.dL_ddist), ptr(self.ws_incl), ptr(self.wts_incl), ptr(self.ws), ptr(self.deltas), ptr(self.ts),
                 ptr(self.rays_a), R, S, ptr(self.dL_dws), st)
        call("mfn_composite_train_bw", ptr(self.dL_dopacity), ptr(self.dL_ddepth), ptr(self.dL_drgb), ptr(self.dL_dws), ptr(self.sigmas), ptr(self.rgbs),
             ptr(self.ws), ptr(self.deltas), ptr(self.ts), ptr(self.rays_a), ptr(self.opacity), ptr(self.depth), ptr(self.rgb), self.T_thr, R, S,
             ptr(self.dL_dsigmas), ptr(self.dL_drgbs), st)

    def _field_back(self):
        """field backward (MLP dgrad / wgrad, hash-grid scatter) into self.grads"""
        st, S = stream_ptr(self.dev), self.cap
        cfg = ctypes.byref(self.cfg)
        n_dev = self._n_back if self._fast_front else ptr(self.n_field)
        if not self._fast_front:          # (the fused compositing kernel of the front stage has cleared the flag)
            self.overflow.zero_()
        if self._fused:
            call("mfn_field_bwd_amp", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(self.xyzs), S, n_dev, ptr(self.dL_dsigmas), ptr(self.dL_drgbs),
                 ptr(self._amp), ptr(self.grads), ptr(self.grads[self.off_rgb:]), ptr(self.overflow), ptr(self.field_ws), self.field_ws.numel(), st)
        else:       # MFN_FIELD_IMPL=v1 (A/B measurements): static loss scale, the AMP state's scale is never changed (see _amp_update)
            call("mfn_field_bwd", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(self.xyzs), S, n_dev, ptr(self.dL_dsigmas), ptr(self.dL_drgbs),
                 self.loss_scale, ptr(self.grads), ptr(self.grads[self.off_rgb:]), ptr(self.overflow), ptr(self.field_ws), self.field_ws.numel(), st)

    def _amp_update(self, st):
        backoff, growth, interval, lo, hi = self._amp_rule if self._fused else (1.0, 1.0, 1 << 30, 1.0, 65536.0)
        call("mfn_amp_update", ptr(self._amp), ptr(self.overflow), backoff, growth, int(interval), lo, hi, 0.9, 0.999, st)

    def _optimizer_step(self, lr=None):
        self.step_count += 1
        lr = float(self.lr if lr is None else lr)
        if lr != self._lr_on_device:
            self._adam_hyper[:1].fill_(lr); self._lr_on_device = lr
        st = stream_ptr(self.dev)
        call("mfn_adam_step_amp", ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.params_h), self.n_params,
             ptr(self._adam_hyper), 0.9, 0.999, 1e-15, 1.0 / self.world_size, ptr(self._amp), ptr(self.overflow), 1, st)
        self._amp_update(st)

    def _adam_from_device_scalars(self, st):
        call("mfn_adam_step_amp", ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.params_h), self.n_params,
             ptr(self._adam_hyper), 0.9, 0.999, 1e-15, 1.0, ptr(self._amp), ptr(self.overflow), 1, st)
        self._amp_update(st)

    def amp_state(self):
        """{loss_scale, skipped_steps, applied_steps} -- a step whose fp16 gradients overflowed is skipped and reported here, never silent"""
        self._wait_comm()
        a = self._amp.tolist()
        return {"loss_scale": a[0], "skipped_steps": int(a[2]), "applied_steps": int(a[3])}

    def _upload_adam_scalars(self, lr, stream):
        """{lr, 1 - beta1^t, 1 - beta2^t} of the step being enqueued -> device memory, through a ring of pinned rows, on `stream` (the back
        stream, where it queues behind the previous step's optimiser and ahead of this step's backward: off the critical path)"""
        if not hasattr(self, "_hyper_host"):
            self._hyper_host = torch.zeros(16, 4).pin_memory()
            self._hyper_evt = [torch.cuda.Event() for _ in range(16)]
            self._hyper_used = [False] * 16
            self._hyper_k = 0
        k = self._hyper_k % 16
        self._hyper_k += 1
        if self._hyper_used[k]:
            self._hyper_evt[k].synchronize()      # 16 steps ago: long done unless the host is that far ahead
        self._lr_on_device = None                 # (the plain path's cached value no longer describes _adam_hyper[0])
        row = self._hyper_host[k]
        _lib.check(_lib.lib.mfn_adam_hyper(float(lr), 0.9, 0.999, int(self.step_count), ctypes.c_void_p(row.data_ptr())), "mfn_adam_hyper")
        with torch.cuda.stream(stream):
            self._adam_hyper.copy_(row, non_blocking=True)
            self._hyper_evt[k].record(stream)
        self._hyper_used[k] = True

    def capture(self):
        """capture _forward_backward into a CUDA graph (call after at least one eager step)"""
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self._forward_backward()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.grads.zero_()
        l0 = _lib.lib.mfn_launch_count()
        if self.dp:
            # three graphs, three stages of the step pipeline (see _step_dp)
            ga, gb, gc = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                self._march()
            with torch.cuda.graph(gb):
                self._field_front()
            self._graph_adam = None
            with torch.cuda.graph(gc):
                self._field_back()
                if not self.collectives and not self._march_after_scatter:      # one GPU: the optimiser is the last node of the back graph (its scalars come from device memory)
                    self._adam_from_device_scalars(stream_ptr(self.dev))
            if not self.collectives and self._march_after_scatter:              # ... or a graph of its own, so that the next step's march can start between the two
                self._graph_adam = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph_adam):
                    self._adam_from_device_scalars(stream_ptr(self.dev))
            self._graph_march, self._graph, self._graph_back = ga, gb, gc
            self._graph_back_only = gc
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._forward_backward()
            self._graph = g
        self.launches_per_forward_backward = int(_lib.lib.mfn_launch_count() - l0)
        if self.dp and not self.collectives and not self._march_after_scatter:      # the backward pass alone, without the optimiser node (replay_forward_backward)
            gd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gd):
                self._field_back()
            self._graph_back_only = gd
        self.grads.zero_()

    def train_step(self, rays_o=None, rays_d=None, target=None, lr=None, global_step=None):
        """one reference training_step (train.py:164-190).  Inputs (device tensors) are copied into the engine's static buffers."""
        if global_step is None:
            global_step = self.step_count
        if self.dp:                # (_step_dp runs the occupancy update itself, on the march stream: exactly one update per 16 steps)
            self._step_dp(lambda: (self.rays_o.copy_(rays_o, non_blocking=True), self.rays_d.copy_(rays_d, non_blocking=True),
                                   self.target.copy_(target, non_blocking=True)) if rays_o is not None else None, lr, global_step)
            self._wait_comm()      # callers of train_step may read the parameters right away; train_step_packed leaves the optimiser in flight
            return
        if global_step % 16 == 0:                                                            # train.py:165-168
            self.update_density_grid(warmup=global_step < 256)
        if rays_o is not None:
            self.rays_o.copy_(rays_o, non_blocking=True); self.rays_d.copy_(rays_d, non_blocking=True); self.target.copy_(target, non_blocking=True)
        self._run_forward_backward()
        self._optimizer_step(lr)

    def replay_forward_backward(self):
        """replays everything capture() recorded, in order, on the current stream (tests; the training steps replay the pieces on
        their own streams)"""
        if self.dp:
            self._graph_march.replay(); self._graph.replay(); self._graph_back_only.replay()
        else:
            self._graph.replay()

    def _run_forward_backward(self):
        self._wait_loss_read()
        if self._graph is not None:
            self._graph.replay()
            self.graph_replays += 1
        else:
            self._forward_backward()

    def train_step_packed(self, batch, lr=None, global_step=None):
        """as train_step, with the step's rays in ONE (3, R, 3) float32 tensor [rays_o | rays_d | target] -- device memory or pinned
        host memory (then this is the step's only host-to-device copy)."""
        if global_step is None:
            global_step = self.step_count
        if self.dp:
            return self._step_dp(lambda: self._load_batch(batch), lr, global_step)
        if global_step % 16 == 0:
            self.update_density_grid(warmup=global_step < 256)
        self.rays.copy_(batch, non_blocking=True)
        self._run_forward_backward()
        self._optimizer_step(lr)

    # ------------------------------------------------------------------------------------------------ resident data set
    def attach_dataset(self, dataset, seed=0):
        """`dataset`: mfnerf_b200.dataset.ResidentDataset.  train_step_resident() then draws its own batches on the device.
        Data-parallel: every rank holds the data set and draws its OWN rays (the reference's per-rank sampling, datasets/base.py:22-44)."""
        rank = torch.distributed.get_rank(self.pg) if self.world_size > 1 else 0
        self.dataset, self._batch_seed = dataset, mdist.shard_seed(seed, rank)

    def _draw_batch(self):
        """one launch: random (image, pixel) per ray -> rays_o, rays_d, target straight into the step's static buffers
        (datasets/base.py:22-34 + train.py:83-96 + ray_utils.py:23-35, 60-68); the marcher's call counter makes every step's draw new"""
        from . import dataset as mds
        ds = self.dataset
        mds.ray_batch(ds.camera, ds.poses, self.n_rays, pixels=ds.pixels, strategy=ds.strategy, seed=self._batch_seed,
                      call_counter=self.march_ws[16:], rays_o=self.rays_o, rays_d=self.rays_d, rgb=self.target)

    def train_step_resident(self, lr=None, global_step=None):
        """as train_step, with the batch drawn on the device from the attached data set: the step has no host-to-device traffic"""
        if global_step is None:
            global_step = self.step_count
        if self.dp:
            return self._step_dp(self._draw_batch, lr, global_step)
        if global_step % 16 == 0:
            self.update_density_grid(warmup=global_step < 256)
        self._draw_batch()
        self._run_forward_backward()
        self._optimizer_step(lr)

    def mark_invisible_cells(self, K, poses, img_wh):
        """networks.py:199-240 (train.py:159-162, once before training): density_grid <- 0 / -1, self.count_grid <- camera coverage"""
        K = torch.as_tensor(K, dtype=torch.float32).to(self.dev).contiguous()
        poses = torch.as_tensor(poses, dtype=torch.float32).to(self.dev).contiguous()
        self.count_grid = torch.zeros_like(self.density_grid)
        call("mfn_grid_mark_invisible", ptr(K), ptr(poses), poses.shape[0], int(img_wh[0]), int(img_wh[1]), self.cascades, self.scale, G, NEAR_DISTANCE,
             ptr(self.density_grid), ptr(self.count_grid), stream_ptr(self.dev))

    # ------------------------------------------------------------------------------------------------ data-parallel step
    def _dp_setup(self):
        """ZeRO-1 style sharding of the optimiser over the ranks: reduce-scatter of the flat fp32 gradient, Adam on this rank's
        1/N slice of (params, exp_avg, exp_avg_sq), all-gather of the fp16 shadow parameters every kernel reads.  Compared with
        all-reduce + replicated Adam this moves 25% fewer bytes over NVLink and divides the optimiser's HBM traffic by N."""
        W = self.world_size
        assert self.n_params % (8 * W) == 0, "flat parameter vector is padded to a multiple of 8 * world_size"
        self._shard = self.n_params // W
        self._rank = torch.distributed.get_rank(self.pg) if self.collectives else 0
        self._grad_shard = torch.zeros(self._shard, device=self.dev) if self.collectives else None
        r, n = self._rank, self._shard
        sl = slice(r * n, (r + 1) * n)
        # static views / pointers of this rank's shard (the step loop is host-bound at several GPUs: no per-step slicing)
        self._sl_ph = self.params_h[sl]
        self._adam_ptrs = (ptr(self.params[sl]), ptr(self._grad_shard if self.collectives else self.grads), ptr(self.exp_avg[sl]), ptr(self.exp_avg_sq[sl]),
                           ptr(self.params_h[sl]))
        self._march_stream = torch.cuda.Stream(self.dev, priority=int(os.environ.get("MFN_MARCH_PRIO", "0")))
        self._march_done = torch.cuda.Event()
        self._cb_done = torch.cuda.Event()
        self._back_done = torch.cuda.Event()
        self._back_pending = False
        # high priority: the back stream (field backward, scatter, Adam) is the step's critical path; the marching front that overlaps it
        # (default priority) only gets the issue slots it leaves free
        self._comm_stream = torch.cuda.Stream(self.dev, priority=int(os.environ.get("MFN_BACK_PRIO", "-1")))
        self._comm_stream_ptr = ctypes.c_void_p(self._comm_stream.cuda_stream)
        self._comm_done = torch.cuda.Event()
        self._comm_pending = False

    def _load_batch(self, batch):
        """(3, R, 3) [rays_o | rays_d | target] -> self.rays.  A HOST batch (pinned memory) crosses PCIe on its own stream into a ring of
        staging buffers -- the host runs ahead of the GPU, so the transfer of step t+k overlaps the kernels of step t -- and the main
        stream only pays a device-to-device copy."""
        if batch.is_cuda:
            self.rays.copy_(batch, non_blocking=True)
            return
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty_like(self.rays) for _ in range(4)]
            self._stage_ready = [torch.cuda.Event() for _ in range(4)]
            self._stage_free = [torch.cuda.Event() for _ in range(4)]
            self._h2d_stream = torch.cuda.Stream(self.dev)
            self._stage_k = 0
            for e in self._stage_free:
                e.record(torch.cuda.current_stream(self.dev))
        k = self._stage_k % 4
        self._stage_k += 1
        main = torch.cuda.current_stream(self.dev)
        self._h2d_stream.wait_event(self._stage_free[k])
        with torch.cuda.stream(self._h2d_stream):
            self._stage[k].copy_(batch, non_blocking=True)
            self._stage_ready[k].record(self._h2d_stream)
        main.wait_event(self._stage_ready[k])
        self.rays.copy_(self._stage[k], non_blocking=True)
        self._stage_free[k].record(main)

    def loss_to_host(self, dst_pinned):
        """asynchronous read-back of the 3 loss terms of the step just enqueued into pinned host memory, on a side stream (the main
        stream is not held up by the PCIe round trip; the next step's loss kernel waits for the read to be done)"""
        if not hasattr(self, "_d2h_stream"):
            self._d2h_stream = torch.cuda.Stream(self.dev)
            self._loss_ready = torch.cuda.Event()
            self._loss_read = torch.cuda.Event()
        self._loss_ready.record(torch.cuda.current_stream(self.dev))
        self._d2h_stream.wait_event(self._loss_ready)
        with torch.cuda.stream(self._d2h_stream):
            dst_pinned.copy_(self.loss_terms, non_blocking=True)
            self._loss_read.record(self._d2h_stream)
        self._loss_read_pending = True

    def _wait_loss_read(self):
        if getattr(self, "_loss_read_pending", False):
            torch.cuda.current_stream(self.dev).wait_event(self._loss_read)
            self._loss_read_pending = False

    def flush(self):
        """the current stream waits for everything the engine still has in flight on its own streams (the last step's backward pass,
        optimiser and loss read-back): call before timing or reading results produced by train_step_packed / train_step_resident"""
        self._wait_comm()
        self._wait_loss_read()

    def _wait_comm(self):
        if getattr(self, "_comm_pending", False):
            torch.cuda.current_stream(self.dev).wait_event(self._comm_done)
            self._comm_pending = False

    def _step_dp(self, copy_batch, lr, global_step):
        """One training step as a three-stage pipeline over three streams (same arithmetic, same order per datum as the plain path):
          march stream : [occupancy update every 16 steps] -> batch copy -> AABB / jitter / ray marching of step t.  It starts as soon
                         as step t-1's compositor backward has released the sample arrays and the ray batch (with the fused field
                         kernels the backward pass reads only its workspace), so it overlaps step t-1's backward, scatter and Adam
          main stream  : field forward, compositing, loss, compositing backward of step t            (CUDA graph)
          back stream  : field backward + hash-grid scatter (CUDA graph), gradient exchange when data-parallel, Adam; only waited for
                         right before step t+1's field forward."""
        if not hasattr(self, "_shard"):
            self._dp_setup()
        if global_step is None:
            global_step = self.step_count
        main = torch.cuda.current_stream(self.dev)
        ms, cs = self._march_stream, self._comm_stream
        ms.wait_stream(main)         # everything enqueued on the main stream so far: step t-1's field front, and whatever the caller did
        if (self.collectives or self._march_after_scatter) and self._back_pending:
            # data-parallel: the gradient exchange leaves the SMs mostly idle, so the marching front is better spent overlapping IT
            # than competing with the backward / scatter kernels for issue slots (measured on 8 GPUs: 0.478 vs 0.507 ms/step)
            ms.wait_event(self._back_done)
            self._back_pending = False
        with torch.cuda.stream(ms):
            if not self._deep:
                self._wait_comm()    # unfused field shapes: the backward still reads the sample arrays
            if global_step % 16 == 0:                    # the occupancy update queries the field: needs the updated parameters --
                plan = self._dg_prepare(warmup=global_step < 256)      # -- but choosing the cells does not: overlaps the previous step's back stage
                self._wait_comm()
                self._dg_apply(plan)
            copy_batch()
            if self._graph is not None:
                self._graph_march.replay()
            else:
                self._march()
            self._march_done.record(ms)
        main.wait_event(self._march_done)
        self._wait_comm()
        self._wait_loss_read()
        if self._graph is not None:
            self._graph.replay()
            self.graph_replays += 1
        else:
            self._field_front()
        self._cb_done.record(main)
        self.step_count += 1
        p_, g_, m_, v_, ph_ = self._adam_ptrs
        lr_ = float(self.lr if lr is None else lr)
        self._upload_adam_scalars(lr_, cs)
        cs.wait_event(self._cb_done)
        with torch.cuda.stream(cs):
            if self._graph is not None:
                self._graph_back.replay()          # one GPU: ends with the optimiser (whole vector, gradient buffer zeroed by it)
                if self._graph_adam is not None:
                    self._back_done.record(cs); self._back_pending = True
                    self._graph_adam.replay()
            else:
                self._field_back()
                if not self.collectives:
                    if self._march_after_scatter:
                        self._back_done.record(cs); self._back_pending = True
                    self._adam_from_device_scalars(self._comm_stream_ptr)
            if self.collectives and self._symm is not None:
                # ONE kernel: sum this rank's shard of every rank's gradients through the NVSwitch, clear it everywhere, Adam on the fp32
                # master shard, fp16 shadow stored into every rank's copy -- bracketed by two device-side barriers over the symmetric allocation
                self._back_done.record(cs); self._back_pending = True
                sy = self._symm
                sy.barrier(0)                  # every rank's field backward + scatter has landed in its gradient buffer
                call("mfn_dp_exchange_adam", self.world_size, sy.grads_ptrs, sy.shadow_ptrs, sy.flag_ptrs, sy.grads_mc, sy.shadow_mc, p_, m_, v_,
                     self._rank * self._shard, self._shard, ptr(self._adam_hyper), 0.9, 0.999, 1e-15, ptr(self._amp), ptr(self._skip), self._comm_stream_ptr)
                call("mfn_amp_update", ptr(self._amp), ptr(self._skip), *self._amp_rule, 0.9, 0.999, self._comm_stream_ptr)
                sy.barrier(1)                  # every rank's shadow stores are visible and every rank has read this rank's gradients
                self._comm_done.record(cs)     # the next step's field forward waits for the barrier, not for the gradient clear behind it:
                self.grads.zero_()             # local HBM (clearing all copies from the shard owner doubles the NVLink store traffic), and
                self._comm_pending = True      # stream-ordered before the next backward pass, the first kernel to touch the buffer again
                return
            elif self.collectives:
                self._back_done.record(cs); self._back_pending = True
                torch.distributed.reduce_scatter_tensor(self._grad_shard, self.grads, op=torch.distributed.ReduceOp.SUM, group=self.pg)
                torch.distributed.all_reduce(self.overflow, op=torch.distributed.ReduceOp.MAX, group=self.pg)
                call("mfn_adam_step_amp", p_, g_, m_, v_, ph_, self._shard, ptr(self._adam_hyper), 0.9, 0.999, 1e-15, 1.0 / self.world_size,
                     ptr(self._amp), ptr(self.overflow), 0, self._comm_stream_ptr)
                self._amp_update(self._comm_stream_ptr)
                self.grads.zero_()
                torch.distributed.all_gather_into_tensor(self.params_h, self._sl_ph, group=self.pg)
            self._comm_done.record(cs)
        self._comm_pending = True

    def gather_master_params(self):
        """world_size > 1: the fp32 master parameters are sharded over the ranks; returns the full vector (state_dict, tests)"""
        self._wait_comm()
        if not self.collectives:
            return self.params
        torch.cuda.current_stream(self.dev).synchronize()
        full = torch.empty_like(self.params)
        r, n = self._rank, self._shard
        torch.distributed.all_gather_into_tensor(full, self.params[r * n:(r + 1) * n].clone(), group=self.pg)
        return full

    def snapshot(self):
        """copy of everything a training step mutates (bench.py restores it so that every timed region sees the same workload)"""
        self._wait_comm()
        return {"params": self.params.clone(), "params_h": self.params_h.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "density_grid": self.density_grid.clone(), "density_bitfield": self.density_bitfield.clone(), "step_count": self.step_count, "_amp": self._amp.clone(),
                "rng": torch.cuda.get_rng_state(self.dev)}

    def restore(self, snap):
        self._wait_comm()
        for k in ("params", "params_h", "exp_avg", "exp_avg_sq", "density_bitfield", "_amp"):
            getattr(self, k).copy_(snap[k])          # in place: a captured graph holds these addresses
        self.density_grid.copy_(snap["density_grid"])
        self.grads.zero_()
        self.step_count = snap["step_count"]
        torch.cuda.set_rng_state(snap["rng"], self.dev)

    def repack_bitfield(self, threshold):
        """density_bitfield <- density_grid > threshold (vren.packbits, raymarching.cu:122-161)"""
        call("mfn_packbits", ptr(self.density_grid), 0, self.density_bitfield.numel(), float(threshold), ptr(self.density_bitfield), stream_ptr(self.dev))

    # ------------------------------------------------------------------------------------------------ test-time rendering
    @torch.no_grad()
    def render(self, rays_o, rays_d, max_samples=MAX_SAMPLES, T_threshold=1e-4, iterations_per_batch=8, min_chunk=None):
        """render(test_time=True) (rendering.py:46-118) for (N,3) rays -> dict(rgb, depth, opacity, total_samples): the device-side
        wavefront of csrc/render.cu.  The host only looks at the alive count between batches of `iterations_per_batch` iterations."""
        d = self.dev
        self._wait_comm()
        if _lib.lib.mfn_field_is_fused(ctypes.byref(self.cfg)) != 1:      # shapes outside the fused kernels (MixedFeature grid, 128-wide rgb net):
            return self.render_reference_loop(rays_o, rays_d, max_samples, T_threshold)      # the same loop, one kernel launch per operation
        N = rays_o.shape[0]
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        min_samples = (1 if self.esf == 0 else 4) if min_chunk is None else int(min_chunk)     # rendering.py:70
        need = _lib.lib.mfn_render_workspace_bytes(N, min_samples)
        if getattr(self, "_render_ws", None) is None or self._render_ws.numel() < need:
            self._render_ws = torch.empty(need, dtype=torch.uint8, device=d)
            self._render_status = torch.zeros(10, dtype=torch.int32).pin_memory()
        ws, status = self._render_ws, self._render_status
        opacity = torch.empty(N, device=d); depth = torch.empty(N, device=d); rgb = torch.empty(N, 3, device=d)
        center = (ctypes.c_float * 3)(0.0, 0.0, 0.0); half = (ctypes.c_float * 3)(self.scale, self.scale, self.scale)
        st = stream_ptr(d)
        call("mfn_render_begin", ptr(rays_o), ptr(rays_d), center, half, N, NEAR_DISTANCE, min_samples, int(max_samples), ptr(opacity), ptr(depth), ptr(rgb), ptr(ws),
             ws.numel(), st)
        cfg = ctypes.byref(self.cfg)
        done_evt = torch.cuda.Event()
        while True:
            call("mfn_render_iterations", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(rays_o), ptr(rays_d), N, ptr(self.density_bitfield),
                 self.cascades, self.scale, self.esf, G, int(max_samples), min_samples, float(T_threshold), int(iterations_per_batch), ptr(opacity), ptr(depth),
                 ptr(rgb), ptr(ws), ws.numel(), st)
            call("mfn_render_status", ptr(ws), status.data_ptr(), st)
            done_evt.record(torch.cuda.current_stream(d))
            done_evt.synchronize()
            s = status.tolist()
            if s[s[2]] == 0 or s[5] >= max_samples:
                break
        call("mfn_render_finish", ptr(rgb), ptr(opacity), self.bg, N, st)
        total = (s[8] & 0xffffffff) | (s[9] << 32)
        return {"rgb": rgb, "depth": depth, "opacity": opacity, "total_samples": total, "iterations": s[6], "field_rows": s[7] * 1024}

    @torch.no_grad()
    def render_reference_loop(self, rays_o, rays_d, max_samples=MAX_SAMPLES, T_threshold=1e-4):
        """render(test_time=True) (rendering.py:46-118) for (N,3) rays -> dict(rgb, depth, opacity, total_samples)"""
        import vren
        d = self.dev
        self._wait_comm()
        N = rays_o.shape[0]
        _, hits_t, _ = vren.ray_aabb_intersect(rays_o, rays_d, self.center, self.half_size, 1)
        hits_t = hits_t[:, 0].contiguous()
        call("mfn_clamp_near", ptr(hits_t), N, NEAR_DISTANCE, stream_ptr(d))
        opacity = torch.zeros(N, device=d); depth = torch.zeros(N, device=d); rgb = torch.zeros(N, 3, device=d)
        alive = torch.arange(N, device=d)
        samples = total = 0
        min_samples = 1 if self.esf == 0 else 4
        cfg = ctypes.byref(self.cfg)
        while samples < max_samples:
            n_alive = alive.shape[0]
            if n_alive == 0:
                break
            ns = max(min(N // n_alive, 64), min_samples)
            samples += ns
            xyzs, dirs, deltas, ts, n_eff = vren.raymarching_test(rays_o, rays_d, hits_t, alive, self.density_bitfield, self.cascades, self.scale,
                                                                  self.esf, G, MAX_SAMPLES, ns)
            total += n_eff.sum()
            n = n_alive * ns
            need = _lib.lib.mfn_field_workspace_bytes(cfg, n, 0)
            if self._cells_ws is None or self._cells_ws.numel() < need:
                self._cells_ws = torch.empty(need, dtype=torch.uint8, device=d)
            sig = torch.empty(n, device=d); col = torch.empty(n, 3, device=d)
            # padded (all-zero) rows are evaluated too and then ignored by the compositor via N_eff -- cheaper than the
            # reference's boolean-mask gather/scatter round trip (rendering.py:91-97) and free of its host syncs
            dirs_v = dirs.view(-1, 3)
            call("mfn_field_fwd", cfg, ptr(self.xyz_params_h), ptr(self.rgb_params_h), ptr(xyzs), ptr(dirs_v), n, None, ptr(sig), ptr(col),
                 ptr(self._cells_ws), self._cells_ws.numel(), stream_ptr(d))
            vren.composite_test_fw(sig.view(n_alive, ns), col.view(n_alive, ns, 3), deltas, ts, hits_t, alive, T_threshold, n_eff, opacity, depth, rgb)
            alive = alive[alive >= 0]
        bg = torch.tensor(list(self.bg), device=d)
        rgb = rgb + bg * (1 - opacity)[:, None]
        return {"rgb": rgb, "depth": depth, "opacity": opacity, "total_samples": total}
