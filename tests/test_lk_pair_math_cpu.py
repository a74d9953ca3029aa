"""CPU suite: the integer identities the LK tracker's iteration loop relies on (csrc/k_lk.cu, K8 (vi) in DESIGN.md), checked in numpy
with the PTX semantics of dp2a restated here.  The kernel multiplies the two gradients of a pixel PAIR with the pair's differences
in IDP.2A on the bytes of the differences; these tests pin that the decomposition gives the exact sums of
calcOpticalFlowPyrLK's b1 = sum(diff * Ix), b2 = sum(diff * Iy) (tracking.cpp:593-618 through cv::calcOpticalFlowPyrLK) and that
every partial sum fits the 32-bit registers and the 32-bit warp reductions it is kept in."""
import numpy as np

DMAX = 8160   # |diff|: interpolated pixels carry 5 fraction bits (255 * 32)
GMAX = 4080   # |Ix|, |Iy|: Scharr of 8-bit pixels (3 + 10 + 3) * 255
WIN = 21


def _s16(x):
    return ((np.asarray(x, np.int64) + 0x8000) & 0xFFFF) - 0x8000


def dp2a(a, b, c, hi, b_signed):
    """dp2a.{lo,hi}.s32.{u32,s32}: d = c + a.h0 * b.byte[2 hi] + a.h1 * b.byte[2 hi + 1]; a = two signed 16-bit halves."""
    a = np.asarray(a, np.int64) & 0xFFFFFFFF
    b = np.asarray(b, np.int64) & 0xFFFFFFFF
    a0, a1 = _s16(a & 0xFFFF), _s16(a >> 16)
    sh = 16 if hi else 0
    b0, b1 = (b >> sh) & 0xFF, (b >> (sh + 8)) & 0xFF
    if b_signed:
        b0, b1 = ((b0 + 0x80) & 0xFF) - 0x80, ((b1 + 0x80) & 0xFF) - 0x80
    return np.asarray(c, np.int64) + a0 * b0 + a1 * b1


def prmt_5140(x, y):
    """__byte_perm(x, y, 0x5140): (x.b0, y.b0, x.b1, y.b1)"""
    x = np.asarray(x, np.int64) & 0xFFFFFFFF
    y = np.asarray(y, np.int64) & 0xFFFFFFFF
    return (x & 0xFF) | ((y & 0xFF) << 8) | (((x >> 8) & 0xFF) << 16) | (((y >> 8) & 0xFF) << 24)


def pack16(lo, hi):
    return (np.asarray(lo, np.int64) & 0xFFFF) | ((np.asarray(hi, np.int64) & 0xFFFF) << 16)


def pair_products(d0, d1, gx0, gx1, gy0, gy1):
    """What one row of the iteration loop adds to (b1l, b1h, b2l, b2h) for the pixel pair (d0, d1)."""
    dd = prmt_5140(d0, d1)
    gxp, gyp = pack16(gx0, gx1), pack16(gy0, gy1)
    return (dp2a(gxp, dd, 0, False, False), dp2a(gxp, dd, 0, True, True), dp2a(gyp, dd, 0, False, False), dp2a(gyp, dd, 0, True, True))


def test_pair_decomposition_is_exact_on_random_and_extreme_values():
    rng = np.random.default_rng(7)
    n = 200000
    d0, d1 = rng.integers(-DMAX, DMAX + 1, n), rng.integers(-DMAX, DMAX + 1, n)
    g = [rng.integers(-GMAX, GMAX + 1, n) for _ in range(4)]
    # the corners of the value ranges and the byte boundaries of the differences
    edge_d = np.array([-DMAX, -DMAX + 1, -256, -255, -1, 0, 1, 127, 128, 255, 256, 257, DMAX - 1, DMAX])
    edge_g = np.array([-GMAX, -1, 0, 1, GMAX])
    ed0, ed1, eg0, eg1 = [m.ravel() for m in np.meshgrid(edge_d, edge_d, edge_g, edge_g)]
    d0, d1 = np.concatenate([d0, ed0]), np.concatenate([d1, ed1])
    gx0, gx1 = np.concatenate([g[0], eg0]), np.concatenate([g[1], eg1])
    gy0, gy1 = np.concatenate([g[2], eg1]), np.concatenate([g[3], eg0])
    l1, h1, l2, h2 = pair_products(d0, d1, gx0, gx1, gy0, gy1)
    assert np.array_equal(l1 + 256 * h1, d0 * gx0 + d1 * gx1)
    assert np.array_equal(l2 + 256 * h2, d0 * gy0 + d1 * gy1)


def test_partial_sums_fit_their_registers_and_reductions():
    # a lane accumulates 7 pairs = 14 pixels; the warp reduction adds the 32 lanes = the 441 window pixels (+ 7 zero entries)
    lane_low = 14 * 255 * GMAX        # |sum of gradient * unsigned low byte|
    lane_high = 14 * 32 * GMAX        # |sum of gradient * signed high byte|: |diff >> 8| <= 32
    win_low = WIN * WIN * 255 * GMAX
    win_high = WIN * WIN * 32 * GMAX
    assert -(-DMAX >> 8) == 32 and DMAX >> 8 == 31
    assert lane_low < 2**31 and lane_high < 2**31
    assert win_low < 2**29 and win_high < 2**26          # what the kernel's comment states
    assert win_low + 256 * win_high > 2**32              # ... while the full sum does NOT fit 32 bits: 64-bit recombination is needed
    # worst case through the actual data path: every pixel at the extremes
    d = np.full(14, -DMAX)
    gx = np.full(14, GMAX)
    l1, h1, _, _ = pair_products(d[0::2], d[1::2], gx[0::2], gx[1::2], gx[0::2], gx[1::2])
    assert abs(int(l1.sum())) <= lane_low and abs(int(h1.sum())) <= lane_high
    assert int(l1.sum()) + 256 * int(h1.sum()) == 14 * -DMAX * GMAX


def test_err_and_template_fields_of_a_pair_entry():
    # entry = {x0, x1, Ix0 | Ix1 << 16, Iy0 | Iy1 << 16}: the 16-bit halves must round-trip signed gradients, and strip 63's zero halves
    # must add nothing whatever its difference is
    rng = np.random.default_rng(11)
    gx0 = rng.integers(-GMAX, GMAX + 1, 1000)
    w = pack16(gx0, 0)
    assert np.array_equal(_s16(w & 0xFFFF), gx0) and np.all((w >> 16) == 0)
    d0, d1 = rng.integers(-DMAX, DMAX + 1, 1000), rng.integers(-DMAX, DMAX + 1, 1000)
    l1, h1, l2, h2 = pair_products(d0, d1, gx0, 0, gx0, 0)
    assert np.array_equal(l1 + 256 * h1, d0 * gx0) and np.array_equal(l2 + 256 * h2, d0 * gx0)
