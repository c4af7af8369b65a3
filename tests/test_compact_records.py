"""The arithmetic fact behind the compact box records of the round-based cubature kernel (csrc/hcubature_rb.cuh, kCptBytes):
a box reached from [-1, 1]^d by successive halvings -- HCubature's  w = (b - a) * 0.5;  upper child a += w, lower child b -= w,
the sequence the full-record kernel and the oracle perform -- has corners  a = -1 + i * 2^(1-l),  b = a + 2^(1-l)  (l halvings along
the axis, i = the path's bits), and up to l = 31 (the kernel hands deeper problems back) both forms are the SAME doubles: every
intermediate is a dyadic rational with at most 32 significant bits.  The kernel's decode is one FMA per axis."""
import math

import numpy as np


def _decode(idx, lev):
    w = math.ldexp(1.0, 1 - lev)                       # __hiloint2double((1024 - lev) << 20, 0)
    a = float(np.float64(idx) * np.float64(w) + np.float64(-1.0))   # fma(idx, w, -1): the product is exact, so is the sum
    return a, a + w


def test_dyadic_corners_equal_successive_halving_up_to_level_31():
    rng = np.random.default_rng(20261020)
    for trial in range(2000):
        a, b = -1.0, 1.0
        idx, lev = 0, 0
        depth = int(rng.integers(1, 32))
        for _ in range(depth):
            w = (b - a) * 0.5
            up = bool(rng.integers(0, 2))
            if up:
                a = a + w                               # child 0: upper half
            else:
                b = b - w                               # child 1: lower half
            idx, lev = 2 * idx + (1 if up else 0), lev + 1
            da, db = _decode(idx, lev)
            assert da == a and db == b, (trial, lev, idx, a, b, da, db)
        assert lev <= 31 and idx < 2 ** 31              # fits the record's 32-bit index and 5-bit level


def test_record_fields_fit_32_bytes():
    """uint4 {i_0..i_3} | uint2 {meta, i_4} | double I: 5 levels of 5 bits + a 3-bit split axis in `meta` for every dimension 2..5"""
    for d in range(2, 6):
        meta = 0
        for k in range(d):
            meta |= 31 << (5 * k)
        meta |= (d - 1) << 28
        assert meta < 2 ** 32 and (meta >> 28) == d - 1 and all(((meta >> (5 * k)) & 31) == 31 for k in range(d))
        assert 4 * 4 + 2 * 4 + 8 == 32


def test_level_32_is_where_the_compact_form_stops():
    """one more halving needs a 33-bit index: the kernel flags the problem instead (cpt_max_level <= 31)"""
    idx = 2 ** 31 - 1
    assert 2 * idx + 1 >= 2 ** 32 - 1 and 2 * (idx + 1) >= 2 ** 32
