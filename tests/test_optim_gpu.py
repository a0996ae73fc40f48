"""the fused optimiser (mfn_adam_step / mfn_adam_step_dev) against a float64 restatement of apex FusedAdam as the reference
configures it (train.py:136: FusedAdam(lr, eps=1e-15), betas (0.9, 0.999), bias correction on, no weight decay), with the AMP
bookkeeping the kernel folds in: gradient unscale, skip on overflow, fp16 shadow refresh, gradient zeroing."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_adam(p, g, m, v, lr, step, grad_scale):
    """float64, with the constants the fp32 kernel sees (betas are floats, 1 - beta is taken in fp32, bias corrections in double)"""
    b1, b2 = np.float32(0.9), np.float32(0.999)
    omb1, omb2 = float(np.float32(1) - b1), float(np.float32(1) - b2)
    bc1 = float(np.float32(1.0 - float(b1) ** step)); bc2 = float(np.float32(1.0 - float(b2) ** step))
    gr = (g * np.float32(grad_scale)).double()                      # the kernel unscales in fp32
    m = float(b1) * m.double() + omb1 * gr
    v = float(b2) * v.double() + omb2 * gr * gr
    p = p.double() - float(np.float32(lr)) * (m / bc1) / (torch.sqrt(v / bc2) + 1e-15)
    return p, m, v


def test_adam_step_matches_fused_adam_restatement():
    from mfnerf_b200._lib import call, ptr, stream_ptr, lib
    n = 100_003                                                      # not a multiple of 4: vector body + scalar tail
    gen = torch.Generator(device="cuda").manual_seed(0)
    p = torch.rand(n, device="cuda", generator=gen) - 0.5
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    ph = torch.zeros(n, dtype=torch.float16, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    scale = 1.0 / 1024.0
    rp, rm, rv = p.double(), m.double(), v.double()
    for step in (1, 2, 3):
        g = torch.randn(n, device="cuda", generator=gen) * 30.0
        g[::7] = 0.0                                                 # hash entries no sample touched this step
        g0 = g.clone()
        p_before = p.clone()
        call("mfn_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(ph), n, 1e-2, 0.9, 0.999, 1e-15, step, scale, ptr(flag), 1, stream_ptr())
        rp, rm, rv = _ref_adam(rp.float(), g0, rm.float(), rv.float(), 1e-2, step, scale)     # from the kernel's own fp32 state of the previous step
        # m = 0.9 m + 0.1 g cancels when the two terms have opposite signs (from step 2 on, ~0.6 % of the entries): the fp32 result then
        # carries the ABSOLUTE rounding error of the terms (|0.1 g scale| <= 0.015 -> ulp 9e-10, two roundings, FMA contraction either
        # way) while the sum itself can be 1000x smaller -- round 1's "exp_avg off by 2e-4 relative in 552 / 100 003 entries" under
        # atol 1e-12.  An fp32 emulation of the kernel's expression on CPU gives max |err| = 1.4e-9 and 550-800 entries beyond rtol 1e-6;
        # atol 5e-9 is 3.5x that worst case.  Away from cancellation rtol 1e-6 (8 ulp) binds.
        torch.testing.assert_close(m.double(), rm, rtol=1e-6, atol=5e-9)
        torch.testing.assert_close(v.double(), rv, rtol=2e-6, atol=1e-14)                    # sums of squares: no cancellation
        torch.testing.assert_close(p.double(), rp, rtol=0, atol=2e-7)                        # one fp32 rounding of |p| <= 0.6 plus the update's
        upd, upd_r = (p.double() - p_before.double()), (rp - p_before.double())
        big = upd_r.abs() > 1e-3
        assert int(big.sum()) > n // 2
        torch.testing.assert_close(upd[big], upd_r[big], rtol=2e-4, atol=0)
        assert torch.equal(ph, p.half()) and int((g != 0).sum()) == 0                          # shadow refreshed, gradient zeroed
        rp, rm, rv = p.double(), m.double(), v.double()
    if True:                                                         # entries that never saw a gradient do not move (eps = 1e-15: 0 / (0 + eps))
        fresh_p = torch.rand(64, device="cuda") ; fp0 = fresh_p.clone()
        z = torch.zeros(64, device="cuda"); zm = torch.zeros(64, device="cuda"); zv = torch.zeros(64, device="cuda")
        call("mfn_adam_step", ptr(fresh_p), ptr(z), ptr(zm), ptr(zv), None, 64, 1e-2, 0.9, 0.999, 1e-15, 1, scale, None, 1, stream_ptr())
        assert torch.equal(fresh_p, fp0) and int((zm != 0).sum()) == 0 and int((zv != 0).sum()) == 0
    # overflow flag set (GradScaler found an inf): the step is skipped, the gradient is still cleared
    flag.fill_(1)
    g = torch.randn(n, device="cuda", generator=gen)
    p0, m0, v0, ph0 = p.clone(), m.clone(), v.clone(), ph.clone()
    call("mfn_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(ph), n, 1e-2, 0.9, 0.999, 1e-15, 4, scale, ptr(flag), 1, stream_ptr())
    assert torch.equal(p, p0) and torch.equal(m, m0) and torch.equal(v, v0) and torch.equal(ph, ph0) and int((g != 0).sum()) == 0
    flag.zero_()
    # the same step with its scalars in device memory (the launch that sits inside the CUDA graph): bit-identical
    g = torch.randn(n, device="cuda", generator=gen) * 30.0
    pa, ma, va, ga, pha = p.clone(), m.clone(), v.clone(), g.clone(), ph.clone()
    call("mfn_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(ph), n, 3e-3, 0.9, 0.999, 1e-15, 4, scale, ptr(flag), 0, stream_ptr())
    h = (ctypes.c_float * 4)()
    assert lib.mfn_adam_hyper(3e-3, 0.9, 0.999, 4, h) == 0
    hd = torch.tensor(list(h), device="cuda")
    call("mfn_adam_step_dev", ptr(pa), ptr(ga), ptr(ma), ptr(va), ptr(pha), n, ptr(hd), 0.9, 0.999, 1e-15, scale, ptr(flag), 0, stream_ptr())
    assert torch.equal(pa, p) and torch.equal(ma, m) and torch.equal(va, v) and torch.equal(pha, ph)
    assert torch.equal(ga, g) and int((g != 0).sum()) > 0            # zero_grad = 0 leaves the gradient alone


def test_amp_state_follows_gradscaler_semantics():
    """mfn_adam_step_amp / mfn_amp_update (round 1 advisor finding: the bias-correction step advanced on skipped steps and a fixed loss
    scale skipped silently for ever): the step with its scalars in the device-side AMP state equals mfn_adam_step with the same numbers;
    an overflow skips the step, halves the scale and does NOT advance Adam's t; `growth_interval` applied steps double the scale."""
    from mfnerf_b200._lib import call, ptr, stream_ptr
    n = 4099
    gen = torch.Generator(device="cuda").manual_seed(3)
    p = torch.rand(n, device="cuda", generator=gen) - 0.5
    m = torch.rand(n, device="cuda", generator=gen) * 1e-2; v = torch.rand(n, device="cuda", generator=gen) * 1e-4
    g = torch.randn(n, device="cuda", generator=gen) * 50.0
    ph = torch.zeros(n, dtype=torch.float16, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    scale, world, t = 256.0, 4, 7
    amp = torch.zeros(8, device="cuda")
    call("mfn_amp_init", ptr(amp), scale, t - 1, 0.9, 0.999, stream_ptr())
    amp[1] = 3.0; amp[2] = 1.0
    lr = torch.tensor([3e-3], device="cuda")
    pa, ma, va, ga, pha = p.clone(), m.clone(), v.clone(), g.clone(), ph.clone()
    call("mfn_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(ph), n, 3e-3, 0.9, 0.999, 1e-15, t, 1.0 / (scale * world), ptr(flag), 1, stream_ptr())
    call("mfn_adam_step_amp", ptr(pa), ptr(ga), ptr(ma), ptr(va), ptr(pha), n, ptr(lr), 0.9, 0.999, 1e-15, 1.0 / world, ptr(amp), ptr(flag), 1, stream_ptr())
    torch.testing.assert_close(ma, m, rtol=1e-6, atol=1e-9); torch.testing.assert_close(va, v, rtol=2e-6, atol=1e-14)   # (1/world)/scale vs 1/(scale*world): 1 ulp
    torch.testing.assert_close(pa, p, rtol=0, atol=2e-7)
    assert int((ga != 0).sum()) == 0
    rule = (0.5, 2.0, 4, 1.0, 1024.0, 0.9, 0.999)
    call("mfn_amp_update", ptr(amp), ptr(flag), *rule, stream_ptr())
    assert amp.tolist()[:4] == [512.0, 0.0, 1.0, float(t)]                    # tracker 3 -> 4 = the interval: the scale grows, t advances
    amp[:4] = torch.tensor([256.0, 2.0, 1.0, 6.0])
    call("mfn_amp_update", ptr(amp), ptr(flag), *rule, stream_ptr())
    assert amp.tolist()[:4] == [256.0, 3.0, 1.0, 7.0]                         # applied step: tracker and t advance
    call("mfn_amp_update", ptr(amp), ptr(flag), *rule, stream_ptr())
    assert amp.tolist()[:4] == [512.0, 0.0, 1.0, 8.0]                         # 4th consecutive applied step: the scale grows
    flag.fill_(1)
    p0, m0, v0 = pa.clone(), ma.clone(), va.clone()
    g2 = torch.full((n,), float("inf"), device="cuda")
    call("mfn_adam_step_amp", ptr(pa), ptr(g2), ptr(ma), ptr(va), ptr(pha), n, ptr(lr), 0.9, 0.999, 1e-15, 1.0 / world, ptr(amp), ptr(flag), 1, stream_ptr())
    call("mfn_amp_update", ptr(amp), ptr(flag), *rule, stream_ptr())
    assert torch.equal(pa, p0) and torch.equal(ma, m0) and torch.equal(va, v0) and int((g2 != 0).sum()) == 0
    assert amp.tolist()[:4] == [256.0, 0.0, 2.0, 8.0]                         # skipped: scale halved, t unchanged, counted
    amp[:4] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    call("mfn_amp_update", ptr(amp), ptr(flag), *rule, stream_ptr())
    assert amp.tolist()[0] == 1.0                                          # clamped at min_scale
    flag.zero_()
    amp[:4] = torch.tensor([1024.0, 3.0, 0.0, 0.0])
    call("mfn_amp_update", ptr(amp), ptr(flag), *rule, stream_ptr())
    assert amp.tolist()[0] == 1024.0                                       # ... and at max_scale


def test_engine_recovers_from_overflow_and_reports_skipped_steps():
    """a loss scale far too large overflows the fp16 gradients: the engine skips those steps (parameters untouched), halves the scale
    until the gradients fit, then trains; skipped steps are counted, never silent"""
    import scenes
    import vren
    from mfnerf_b200.engine import NGPEngine
    eng = NGPEngine(scale=0.5, n_rays=256, sample_capacity=256 * 160, log2_T=15, loss_scale=2.0 ** 40)
    eng.density_grid.copy_(torch.from_numpy(scenes.syn.lego_density_grid(0.5, 1)).cuda())
    vren.packbits(eng.density_grid.reshape(-1), 0.5, eng.density_bitfield)
    rays = scenes.scene("lego", 256, seed=3)
    batch = torch.stack([torch.from_numpy(rays["rays_o"]), torch.from_numpy(rays["rays_d"]),
                         scenes.syn.analytic_render(rays["rays_o"], rays["rays_d"]).float()]).cuda().contiguous()
    p0 = eng.gather_master_params().clone()
    eng.train_step_packed(batch, global_step=1)
    st = eng.amp_state()
    assert st["skipped_steps"] == 1 and st["applied_steps"] == 0 and st["loss_scale"] == 2.0 ** 39
    assert torch.equal(eng.gather_master_params(), p0)
    first = None
    for s in range(2, 80):
        if s == 6:
            eng.capture()
        eng.train_step_packed(batch, global_step=s)
        if s == 40:
            eng.flush(); first = float(eng.loss_terms.sum())
    eng.flush()
    st = eng.amp_state()
    assert 5 < st["skipped_steps"] < 40 and st["applied_steps"] == 79 - st["skipped_steps"] and st["loss_scale"] < 2.0 ** 34
    assert float(eng.loss_terms.sum()) < first


def test_fused_exchange_kernel_emulated_ranks():
    """mfn_dp_exchange_adam (csrc/dp_exchange.cu), peer-pointer path, with the ranks EMULATED on one GPU (the guide's rule when there are
    fewer GPUs than ranks): world = 3 gradient / shadow / flag buffers in local memory, one launch per rank's shard.  Expected: every shard
    of every buffer summed and Adam-stepped exactly like mfn_adam_step_amp on the pre-summed gradient; all shadows identical; the
    overflow flag of ANY rank skips the step everywhere.  (The multicast path needs a multi-GPU box:
    tools/dp_exchange_check.py under torchrun.)"""
    import ctypes
    from mfnerf_b200._lib import call, ptr, stream_ptr
    world, n = 3, 3 * 8 * 1000
    shard = n // world
    gen = torch.Generator(device="cuda").manual_seed(5)
    p = torch.rand(n, device="cuda", generator=gen) - 0.5
    m = torch.rand(n, device="cuda", generator=gen) * 1e-2; v = torch.rand(n, device="cuda", generator=gen) * 1e-4
    grads = [torch.randn(n, device="cuda", generator=gen) * 40.0 for _ in range(world)]
    shadows = [torch.zeros(n, dtype=torch.float16, device="cuda") for _ in range(world)]
    flags = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    amp = torch.zeros(8, device="cuda"); call("mfn_amp_init", ptr(amp), 512.0, 4, 0.9, 0.999, stream_ptr())
    lr = torch.tensor([5e-3], device="cuda")
    arr = lambda ts: (ctypes.c_uint64 * world)(*[t.data_ptr() for t in ts])
    # reference: the plain sharded path -- sum, then mfn_adam_step_amp
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    gsum = torch.stack(grads).sum(0)            # (fp32 sum in rank order, like the kernel's loop)
    gs = grads[0].clone()
    for g in grads[1:]:
        gs += g
    shr = torch.zeros(n, dtype=torch.float16, device="cuda")
    call("mfn_adam_step_amp", ptr(pr), ptr(gs), ptr(mr), ptr(vr), ptr(shr), n, ptr(lr), 0.9, 0.999, 1e-15, 1.0 / world, ptr(amp), None, 0, stream_ptr())
    skip = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    for r in range(world):      # "rank r" steps its shard
        sl = slice(r * shard, (r + 1) * shard)
        call("mfn_dp_exchange_adam", world, arr(grads), arr(shadows), arr(flags), 0, 0, ptr(p[sl]), ptr(m[sl]), ptr(v[sl]), r * shard, shard,
             ptr(lr), 0.9, 0.999, 1e-15, ptr(amp), ptr(skip), stream_ptr())
    torch.cuda.synchronize()
    assert int(skip) == 0
    torch.testing.assert_close(m, mr, rtol=1e-6, atol=1e-9); torch.testing.assert_close(v, vr, rtol=2e-6, atol=1e-14)
    torch.testing.assert_close(p, pr, rtol=0, atol=2e-7)
    for s in shadows:
        assert torch.equal(s, p.half())
    del gsum
    # any rank's overflow flag: nothing moves anywhere, the decision is published
    for g in grads:
        g.normal_(generator=gen)
    flags[2].fill_(1)
    p0, m0, v0, s0 = p.clone(), m.clone(), v.clone(), shadows[0].clone()
    for r in range(world):
        sl = slice(r * shard, (r + 1) * shard)
        call("mfn_dp_exchange_adam", world, arr(grads), arr(shadows), arr(flags), 0, 0, ptr(p[sl]), ptr(m[sl]), ptr(v[sl]), r * shard, shard,
             ptr(lr), 0.9, 0.999, 1e-15, ptr(amp), ptr(skip), stream_ptr())
    torch.cuda.synchronize()
    assert int(skip) == 1 and torch.equal(p, p0) and torch.equal(m, m0) and torch.equal(v, v0) and torch.equal(shadows[1], s0)
    # argument checks
    from mfnerf_b200._lib import lib
    assert lib.mfn_dp_exchange_adam(world, arr(grads), arr(shadows), arr(flags), 0, 0, ptr(p), ptr(m), ptr(v), 4, shard, ptr(lr), 0.9, 0.999, 1e-15, ptr(amp), None, None) == -2
    assert lib.mfn_dp_exchange_adam(17, arr(grads), arr(shadows), arr(flags), 0, 0, ptr(p), ptr(m), ptr(v), 0, shard, ptr(lr), 0.9, 0.999, 1e-15, ptr(amp), None, None) == -2
