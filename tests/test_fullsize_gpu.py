"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle is too slow there):
8192 rays / T = 2^19 training batches and 800x800 = 640 000-ray renders of the bench workload."""
import ctypes
import os

import numpy as np
import pytest
import torch

import scenes

pytestmark = pytest.mark.gpu

R = 8192


def _bench_engine(**kw):
    from mfnerf_b200.engine import NGPEngine
    eng = NGPEngine(scale=0.5, n_rays=R, seed=1337, **kw)
    eng.density_grid.copy_(torch.from_numpy(scenes.syn.lego_density_grid(0.5, 1)).cuda())
    eng.repack_bitfield(0.5)
    return eng


def _batch(seed):
    o, d, _, _ = scenes.syn.random_rays(R, seed=seed)
    tgt = scenes.syn.analytic_render(o, d)
    return torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), tgt.cuda().float()


def test_march_full_batch_invariants_and_reference_kernels(ref_vren):
    """8192 rays of the bench scene: prefix-sum layout, monotone ts, constant step, every sample inside an occupied cell; and, when the
    reference's own kernels are built, bit-exact N_samples / ts / xyzs against them at this size"""
    import vren
    sc = scenes.scene("lego", R, seed=3)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    o, d, bits = T(sc["rays_o"]), T(sc["rays_d"]), T(sc["bitfield"])
    _, ht, _ = vren.ray_aabb_intersect(o, d, T(sc["center"]), T(sc["half"]), 1)
    h = T(scenes.near_clamp(ht.cpu().numpy()))
    ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(o, d, h, bits, 1, 0.5, 0.0, T(sc["noise"]), 128, 1024)
    n = int(counter[0])
    assert n == int(ra[:, 2].sum()) == xyzs.shape[0] and n > 50 * R // 4
    assert (ra[:, 0] == torch.arange(R, device="cuda")).all()
    start = torch.cumsum(ra[:, 2], 0) - ra[:, 2]
    assert (ra[:, 1] == start).all()
    assert int(ra[:, 2].max()) <= 1024
    seg = torch.repeat_interleave(torch.arange(R, device="cuda"), ra[:, 2])
    same_ray = seg[1:] == seg[:-1]
    assert (ts[1:][same_ray] > ts[:-1][same_ray]).all()                       # strictly increasing along each ray
    assert (deltas == np.float32(np.float32(1.73205080757) / np.float32(1024))).all()   # exp_step_factor 0: dt = sqrt(3)/max_samples
    torch.testing.assert_close(xyzs, o[seg] + d[seg] * ts[:, None], rtol=0, atol=1e-6)
    assert (dirs == d[seg]).all()
    # occupancy of the cell of every sample (raymarching.cu:211-220 for one cascade: mip 0, bound 0.5)
    cell = torch.clamp(0.5 * (xyzs / 0.5 + 1) * 128, 0, 127).int()
    idx = vren.morton3D(cell.contiguous()).long()
    occ = (bits[idx >> 3] >> (idx & 7).to(torch.uint8)) & 1
    assert int(occ.sum()) == n
    if ref_vren is not None:
        rra, rx, rd, rdl, rts, rc = ref_vren.raymarching_train(o, d, h, bits, 1, 0.5, 0.0, T(sc["noise"]), 128, 1024)
        rra = rra.cpu().numpy(); order = np.argsort(rra[:, 0], kind="stable"); rra = rra[order]
        assert (rra[:, 2] == ra[:, 2].cpu().numpy()).all()                       # sample counts per ray: exact
        sel = np.concatenate([np.arange(s, s + k) for _, s, k in rra]) if n else np.zeros(0, np.int64)
        sel = torch.from_numpy(sel).cuda()
        assert torch.equal(rts[:int(rc[0])][sel], ts) and torch.equal(rx[:int(rc[0])][sel], xyzs)


def test_composite_full_batch_identities():
    """sum(ws) = opacity, 1 - opacity = prod(1 - alpha) up to the termination point, depth = sum(w t), rgb = sum(w c)"""
    import vren
    sc = scenes.scene("lego", R, seed=4)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    o, d, bits = T(sc["rays_o"]), T(sc["rays_d"]), T(sc["bitfield"])
    _, ht, _ = vren.ray_aabb_intersect(o, d, T(sc["center"]), T(sc["half"]), 1)
    h = T(scenes.near_clamp(ht.cpu().numpy()))
    ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(o, d, h, bits, 1, 0.5, 0.0, T(sc["noise"]), 128, 1024)
    n = int(counter[0])
    sig, rgbs = scenes.field_values(n, seed=5)
    sig, rgbs = T(sig), T(rgbs)
    total, opacity, depth, rgb, ws = vren.composite_train_fw(sig, rgbs, deltas, ts, ra, 1e-4)
    seg = torch.repeat_interleave(torch.arange(R, device="cuda"), ra[:, 2])
    z = torch.zeros(R, device="cuda", dtype=torch.float64)
    torch.testing.assert_close(z.index_add(0, seg, ws.double()).float(), opacity, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(z.index_add(0, seg, (ws * ts).double()).float(), depth, rtol=1e-5, atol=1e-6)
    for k in range(3):
        torch.testing.assert_close(z.index_add(0, seg, (ws * rgbs[:, k]).double()).float(), rgb[:, k], rtol=1e-5, atol=1e-6)
    # transmittance: log(1 - opacity) = sum log(1 - alpha) over the samples that received weight (plus the terminating one)
    alpha = 1 - torch.exp(-sig.double() * deltas.double())
    counted = torch.zeros(n, dtype=torch.bool, device="cuda")
    rank = torch.arange(n, device="cuda") - ra[seg, 1]
    counted = rank <= total[seg]                      # total_samples = index of the terminating sample, or N (volumerendering.cu:41-44)
    logT = z.index_add(0, seg[counted], torch.log1p(-alpha[counted].clamp(max=1 - 1e-12)))
    alive = total == ra[:, 2]                          # rays that never hit the threshold: exact identity
    torch.testing.assert_close(torch.exp(logT[alive]).float(), (1 - opacity)[alive], rtol=2e-4, atol=2e-6)
    assert ((1 - opacity)[~alive] <= 1e-4 * 1.001).all()      # terminated rays: T <= T_threshold
    assert (ws[rank > total[seg]] == 0).all()


def test_field_full_size_constant_table_and_subset_oracle():
    """T = 2^19, ~500 k samples through the fused tcgen05 kernels: (1) a constant table encodes every position to that constant
    (the 8 trilinear weights sum to 1), so sigma / rgb depend on the direction only; (2) a random 4096-sample subset against the torch
    restatement (the oracle is too slow for the whole batch)"""
    from oracle import field_ref as fr
    eng = _bench_engine()
    o, d, tgt = _batch(31)
    eng.train_step(o, d, tgt, global_step=1)
    n = int(eng.counter[0])
    assert n > 100_000
    xyzs, dirs = eng.xyzs[:n].clone(), eng.dirs[:n].clone()
    sig, rgb = eng.field(xyzs, dirs)
    sel = torch.randperm(n, device="cuda", generator=torch.Generator("cuda").manual_seed(1))[:4096]
    ref = fr.NGPRef(0.5, log2_T=19, params=(eng.params[:eng.n_xyz].cpu(), eng.params[eng.off_rgb:eng.off_rgb + eng.n_rgb].cpu())).cuda()
    with torch.no_grad():
        sig_r, rgb_r = ref(xyzs[sel], dirs[sel])
    torch.testing.assert_close(sig[sel], sig_r, rtol=3e-2, atol=1e-3)      # fp16 activations vs the fp32 restatement (DESIGN.md tolerances)
    torch.testing.assert_close(rgb[sel], rgb_r, rtol=2e-2, atol=3e-3)
    # constant table
    saved = eng.params_h.clone()
    eng.params_h[eng.n_mlp1:eng.n_xyz] = 0.0078125                           # exactly representable in fp16
    sig_c, rgb_c = eng.field(xyzs, dirs)
    assert float((sig_c - sig_c[0]).abs().max()) <= 2e-3 * float(sig_c[0].abs())       # fp32 weights sum to 1 within rounding, one fp16 rounding
    same_dir = (dirs == dirs[0]).all(1)
    assert int(same_dir.sum()) > 1 and float((rgb_c[same_dir] - rgb_c[0]).abs().max()) <= 2e-3
    eng.params_h.copy_(saved)


def test_backward_full_size_matches_directional_finite_difference():
    """d loss / d (uniform shift of every entry of one level) = sum of that level's gradient entries.  The left side comes from two
    forward passes of the full batch, the right side from the fused backward + scatter kernels: a size-independent check of the
    whole backward path at 8192 rays / 2^19 table entries per level"""
    from mfnerf_b200 import field_ops
    eng = _bench_engine(pipelined=False)
    o, d, tgt = _batch(32)
    noise = torch.rand(R, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    eng.fixed_noise = noise
    with torch.no_grad():                      # a livelier field than the 1e-4 init, so that gradients are well above fp16 noise
        eng.params[eng.n_mlp1:eng.n_xyz].uniform_(-0.2, 0.2, generator=torch.Generator("cuda").manual_seed(4))
        eng.params_h.copy_(eng.params)
    eng.rays_o.copy_(o); eng.rays_d.copy_(d); eng.target.copy_(tgt)
    eng.grads.zero_()
    eng._forward_backward()
    torch.cuda.synchronize()
    assert int(eng.overflow[0]) == 0
    g = (eng.grads[eng.n_mlp1:eng.n_xyz] / eng.loss_scale).double().view(-1, 2)
    loss0 = float(eng.loss_terms.sum())
    _, off, _, _ = field_ops.grid_layout(eng.cfg.grid)
    table = eng.params_h[eng.n_mlp1:eng.n_xyz].view(-1, 2)
    eps = 2.0 ** -7                             # fp16-exact shift, large against the table's fp16 spacing
    for level, feat in ((0, 0), (5, 1), (10, 0), (15, 1)):
        lo, hi = off[level], off[level + 1]
        losses = []
        for sgn in (+1, -1):
            saved = table[lo:hi, feat].clone()
            table[lo:hi, feat] = (saved.float() + sgn * eps).half()
            eng._forward_backward()
            losses.append(float(eng.loss_terms.sum()))
            table[lo:hi, feat] = saved
        fd = (losses[0] - losses[1]) / (2 * eps)
        an = float(g[lo:hi, feat].sum())
        assert abs(fd - an) <= 0.08 * max(abs(fd), abs(an)) + 2e-4, (level, feat, fd, an, loss0)


def test_render_full_frame_wavefront_equals_per_op_loop():
    """800 x 800 = 640 000 rays: the device-side wavefront against the same loop driven op by op"""
    eng = _bench_engine()
    o, d, tgt = _batch(33)
    for s in range(1, 40):
        eng.train_step(o, d, tgt, global_step=s)
    pose = scenes.syn.camera_poses(1, seed=5)[0]
    ro, rd = scenes.syn.image_rays(pose)
    ro, rd = torch.from_numpy(ro).cuda(), torch.from_numpy(rd).cuda()
    a = eng.render(ro, rd)
    b = eng.render_reference_loop(ro, rd)
    assert int(a["total_samples"]) == int(b["total_samples"]) > 640_000
    for k in ("opacity", "depth", "rgb"):
        torch.testing.assert_close(a[k], b[k], rtol=0, atol=1e-6)
    # A coarser schedule (at least 8 samples per ray and iteration) changes the iteration count, not the per-sample arithmetic; what it
    # can change is WHERE the reference's sample budget (rendering.py:69, `while samples < max_samples`, samples += N_samples) cuts off the
    # few rays that are still alive at the end: their last low-weight samples -- hence 1e-5 here, against 1e-6 for the reference schedule
    c = eng.render(ro, rd, min_chunk=8)
    for k in ("opacity", "depth", "rgb"):
        torch.testing.assert_close(c[k], b[k], rtol=0, atol=1e-5)
        assert int(((c[k] - b[k]).abs() > 1e-6).sum()) <= 64
    assert c["iterations"] < a["iterations"]
    assert float(a["opacity"].min()) >= 0 and float(a["opacity"].max()) <= 1 + 1e-5


@pytest.mark.parametrize("name,n_rays", [("lego", 65536), ("full", 8192), ("unbounded", 32768), ("axis", 8192)])
def test_fast_marcher_equals_the_exact_replay(name, n_rays, monkeypatch):
    """march.cuh: march_ray_warp_fast takes every occupied lattice point after PROVING that the reference's sequential visit order
    cannot have skipped one, and hands the ray to the exact replay (march_ray_warp) when the proof fails.  Both must give the same
    bits; the number of re-marched rays is read from the workspace header."""
    import vren
    sc = scenes.scene(name, n_rays, seed=11)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    o, d, bits = T(sc["rays_o"]), T(sc["rays_d"]), T(sc["bitfield"])
    _, ht, _ = vren.ray_aabb_intersect(o, d, T(sc["center"]), T(sc["half"]), 1)
    h = T(scenes.near_clamp(ht.cpu().numpy()))
    args = (o, d, h, bits, sc["cascades"], sc["scale"], sc["esf"], T(sc["noise"]), 128, 1024)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MFN_MARCH_FAST", mode)
        vren.raymarching_train(*args)                       # (allocates / sizes the shim's workspace)
        ws = vren._ws_cache[torch.cuda.current_device()]
        ws[:256].zero_()
        out[mode] = vren.raymarching_train(*args)
        torch.cuda.synchronize()
        out[mode + "_remarched"] = int(ws[:256].view(torch.int64)[3])
    assert out["0_remarched"] == 0
    for a, b in zip(out["0"], out["1"]):
        assert a.shape == b.shape and torch.equal(a, b)
    print(f"{name}: {int(out['1'][5][0])} samples, {out['1_remarched']} of {n_rays} rays re-marched by the exact replay")
    assert out["1_remarched"] < n_rays // 20
