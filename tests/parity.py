"""Shared parity criterion for parameter / input gradients of the field kernels (tests only)."""


def grad_close(got, want, name, big_frac=5e-2, rtol=8e-2, atol_frac=2e-3, max_frac=3e-2, flip_frac=3e-2, min_big=20):
    """every entry within max_frac of the largest one; entries above big_frac of the largest also within rtol relative (+ atol_frac of
    the largest) -- except a bounded share of them (flip_frac).  Why the exception: a hidden unit whose pre-activation is within rounding
    of zero gets ReLU mask 1 in one implementation and 0 in the other (fp32 accumulation order of the tensor cores vs torch's GEMM); that
    one sample's whole contribution to the unit's weight-gradient row then differs, which is ~1/sqrt(N) = 1-2 % of a row summed over
    N = 3000 samples.  tools/bwd_check.py (12 seeds x shapes, fused tcgen05 kernels AND the unfused mma.sync pipeline): typical
    err / bound 0.02, and 2-3 runs in 12 with one such row at 0.5-1.6 % of the largest entry."""
    sc = want.abs().max().item()
    err = (got - want).abs()
    big = want.abs() > big_frac * sc
    assert big.sum() > min_big, name
    viol = err[big] > rtol * want.abs()[big] + atol_frac * sc
    assert viol.float().mean().item() <= flip_frac, (name, int(viol.sum()), int(big.sum()), err[big].max().item(), sc)
    assert err.max().item() <= max_frac * sc, (name, err.max().item(), sc)
