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
