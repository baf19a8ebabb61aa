"""CPU checks of the split-precision arithmetic the CUDA kernels rely on (float64 emulation of the fp16 planes and the
fp32-accumulated tensor-core products; no GPU, no library call)."""
import torch

S = 2.0 ** -6          # kPlaneScale in csrc/kernels.cuh


def f16(x):
    return x.to(torch.float16).to(torch.float64)


SUBNORMAL_HALF_ULP = 2.0 ** -25     # the second plane of a small value is an fp16 subnormal (spacing 2^-24): an ABSOLUTE floor


def test_hi_lo_planes_carry_22_bits():
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(4096, generator=g, dtype=torch.float64) * 3).float().double()
    hi = f16(x)
    lo = f16(x - hi)
    assert ((hi + lo - x).abs() <= 2.0 ** -21 * x.abs() + SUBNORMAL_HALF_ULP).all()


def test_scaled_planes_sum_to_the_value_within_2_pow_minus_17():
    """x1 = fp16((1 - s) x), x2 = fp16(x - x1): the planes every consumer (2-product projection, recurrence, head) sees."""
    g = torch.Generator().manual_seed(1)
    x = torch.tanh(torch.randn(8192, generator=g, dtype=torch.float64)).float().double()      # h-like values in (-1, 1)
    x1 = f16((1 - S) * x)
    x2 = f16(x - x1)
    assert ((x1 + x2 - x).abs() <= 2.0 ** -17 * x.abs() + SUBNORMAL_HALF_ULP).all()


def test_two_product_projection_error_bound():
    """x1.W_hi + x2.W' with W' = fp16(W_hi + W_lo / s) against the exact product: error <= 2^-16 sum |x||W| (3-product split:
    2^-20; W as a single fp16 plane: 2^-12) -- the claim behind gemm_xg2_kernel / gemm_xg_pair_kernel."""
    g = torch.Generator().manual_seed(2)
    x = torch.tanh(torch.randn(256, 256, generator=g, dtype=torch.float64)).float().double()
    w = (torch.randn(512, 256, generator=g, dtype=torch.float64) * 0.09).float().double()
    exact = x @ w.t()
    scale = x.abs() @ w.abs().t()
    wh = f16(w)
    wl = f16(w - wh)
    # three products on hi / lo planes
    xh = f16(x)
    xl = f16(x - xh)
    three = xl @ wh.t() + xh @ wl.t() + xh @ wh.t()
    # two products on scaled planes
    x1 = f16((1 - S) * x)
    x2 = f16(x - x1)
    wp = f16(wh + wl / S)
    two = x1 @ wh.t() + x2 @ wp.t()
    single = xh @ wh.t() + xl @ wh.t()
    e3 = ((three - exact).abs() / scale).max().item()
    e2 = ((two - exact).abs() / scale).max().item()
    e1 = ((single - exact).abs() / scale).max().item()
    assert e3 <= 2.0 ** -20 and e2 <= 2.0 ** -16 and e1 >= 2.0 ** -15, (e3, e2, e1)
    # the head's 3-product kernel fed with scaled planes loses only x2.W_lo
    head = x2 @ wh.t() + x1 @ wl.t() + x1 @ wh.t()
    assert ((head - exact).abs() / scale).max().item() <= 2.0 ** -16


def test_median_as_sliding_popcount():
    """scipy.signal.medfilt on 0 / 1 data with zero padding == (window sum >= (k + 1) / 2): what threshold_median_kernel computes."""
    import numpy as np
    from scipy.signal import medfilt
    rng = np.random.default_rng(3)
    for k in (49, 25, 3):
        for n in (10, 49, 700):
            b = (rng.random(n) < 0.5).astype(np.int64)
            pad = np.concatenate([np.zeros(k // 2, np.int64), b, np.zeros(k // 2, np.int64)])
            win = np.convolve(pad, np.ones(k, np.int64), mode="valid")
            assert np.array_equal((win >= (k + 1) // 2).astype(np.int64), medfilt(b, kernel_size=k).astype(np.int64)), (k, n)
