"""The 16-bit operand format of the persistent rk4 forward kernel and the Euler-Maruyama drift (csrc/stage_tc.cuh, "16-bit
operand format"), restated in numpy: x s = xh + xl / 2048 with xh = fp16(x s), xl = fp16((x s - xh) 2048), three products
wh.rh + 2^-11 (wl.rh + wh.rl) with float32 accumulation.  What the kernels rely on, checked without a GPU:

  * the pair reconstructs x to 2^-22 of its magnitude wherever x s is in FP16's normal range, and to an ABSOLUTE error that
    is negligible next to the largest operand below it (subnormal high parts);
  * the weight scale the device picks (the power of two that puts max|W| into [2^13, 2^14)) keeps both planes finite;
  * nothing overflows below the +-6e4 bound at which the kernels raise their flag (and the bound is below FP16's maximum);
  * the three-product sum matches the float64 product to the accuracy of the TF32 split it replaces.
"""
import numpy as np
import pytest

F16_LIMIT = 6.0e4          # kF16Limit in stage_tc.cuh


def split2(x):
    x = np.asarray(x, dtype=np.float32)
    h = x.astype(np.float16)
    l = ((x - h.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return h, l


def join2(h, l):
    return h.astype(np.float64) + l.astype(np.float64) / 2048.0


def weight_scale(w):
    m = float(np.abs(w).max())
    if m == 0.0:
        return 1.0
    _, ex = np.frexp(m)                    # m = f 2^ex, f in [0.5, 1): the kernel's frexpf
    return float(np.ldexp(1.0, 14 - ex))


def test_pair_reconstructs_to_22_bits_in_the_normal_range():
    rng = np.random.default_rng(0)
    x = (rng.uniform(1.0, 2.0, 200000) * 2.0 ** rng.integers(-13, 15, 200000)).astype(np.float32) * rng.choice([-1, 1], 200000)
    x = x[np.abs(x) <= F16_LIMIT]
    h, l = split2(x)
    assert np.isfinite(h.astype(np.float32)).all() and np.isfinite(l.astype(np.float32)).all()
    rel = np.abs(join2(h, l) - x.astype(np.float64)) / np.abs(x)
    assert rel.max() <= 2.0 ** -22


def test_small_values_keep_a_negligible_absolute_error():
    x = np.float32(10.0) ** np.linspace(-12, -4.3, 4000, dtype=np.float32)       # below FP16's normal range (6.1e-5)
    h, l = split2(x)
    err = np.abs(join2(h, l) - x.astype(np.float64))
    assert err.max() <= 2.0 ** -25 / 2048 * 1.01 + 2.0 ** -36                   # half a subnormal step of the scaled low part


@pytest.mark.parametrize("wmax", [3.7e-4, 0.9, 17.0, 5.0e3, 2.5e7])
def test_device_weight_scale_puts_the_largest_weight_below_2_to_14(wmax):
    rng = np.random.default_rng(1)
    w = (rng.standard_normal(4096) * 0.2).astype(np.float32)
    w *= np.float32(wmax / np.abs(w).max())
    s = weight_scale(w)
    assert np.log2(s) == np.round(np.log2(s))                                   # a power of two: scaling is exact
    top = float(np.abs(w * np.float32(s)).max())
    assert 2.0 ** 13 <= top < 2.0 ** 14
    h, l = split2(w * np.float32(s))
    assert np.isfinite(h.astype(np.float32)).all() and np.isfinite(l.astype(np.float32)).all()
    assert float(np.abs(l.astype(np.float32)).max()) <= 2.0 ** 14               # |xl| <= 2^-11 |x s| 2^11 (+ rounding)


def test_the_overflow_bound_is_inside_the_format():
    assert F16_LIMIT < float(np.finfo(np.float16).max)
    h, l = split2(np.array([F16_LIMIT, -F16_LIMIT], dtype=np.float32))
    assert np.isfinite(h.astype(np.float32)).all() and np.isfinite(l.astype(np.float32)).all()
    with np.errstate(over="ignore"):
        h, _ = split2(np.array([7.0e4], dtype=np.float32))                      # what the flag exists for
    assert not np.isfinite(h.astype(np.float32)).all()


def test_three_products_match_the_float64_contraction_like_the_tf32_split():
    rng = np.random.default_rng(2)
    N, K, B = 64, 577, 48
    W = (rng.standard_normal((N, K)) * 0.3).astype(np.float32)
    R = np.abs(rng.standard_normal((B, K)) * 20.0).astype(np.float32)           # rates: non-negative, tens of Hz
    R[:, -1] = 1.0                                                              # the constant-one column
    s = np.float32(weight_scale(W))
    wh, wl = split2(W * s)
    rh, rl = split2(R)
    f = lambda a: a.astype(np.float32)
    main = f(wh) @ f(rh).T                                                       # float32 accumulation, like tensor memory
    cross = f(wl) @ f(rh).T + f(wh) @ f(rl).T
    got = (main + cross * np.float32(2.0 ** -11)) / s
    ref = W.astype(np.float64) @ R.astype(np.float64).T
    scale = np.abs(W).astype(np.float64) @ np.abs(R).astype(np.float64).T       # sum |w| |r|
    err16 = float((np.abs(got - ref) / scale).max())

    def tf32(x):                                                                 # round to nearest, ties away (cvt.rna.tf32)
        b = x.astype(np.float32).view(np.uint32)
        return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    Wh, Rh = tf32(W), tf32(R)
    Wl, Rl = tf32(W - Wh), tf32(R - Rh)
    got32 = Wh @ Rh.T + (Wl @ Rh.T + Wh @ Rl.T)
    err32 = float((np.abs(got32 - ref) / scale).max())
    assert err16 <= 3e-7 and err16 <= 2.0 * err32 + 1e-8, (err16, err32)
