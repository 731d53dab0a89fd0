"""Pins the oracle's arithmetic (oracle/zb_oracle.c) against an exact-rational emulation of the
skylake-16 order and against hand-derived known answers.  The reference has no tests of its own
(SURVEY.md section 4), so these KATs are derived from /root/reference/src/database/index/lsh.rs:39-43,
:222-225 and /root/reference/src/distance.rs:19-49,103-114."""
import struct

import numpy as np
import pytest

import pyref
from oracle import zb_oracle as zo


def f64bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


@pytest.mark.parametrize("n", [4, 16, 17, 40, 384])
def test_dot_l2sq_match_exact_emulation(n):
    rng = np.random.default_rng(n)
    for _ in range(3):
        a = rng.standard_normal(n).astype(np.float32)
        b = rng.standard_normal(n).astype(np.float32)
        assert np.float32(zo.dot(a, b)) == pyref.dot16(a, b)
        assert np.float32(zo.l2sq(a, b)) == pyref.l2sq16(a, b)


def test_order_sensitive_case_distinguishes_lane_order():
    # lane 0 receives 2^24, then +1 three times (each absorbed: 2^24+1 rounds to 2^24 in f32), lane 1..15
    # receive the rest.  A sequential (single accumulator) sum, or a pairwise one, gives a different value.
    n = 64
    a = np.ones(n, dtype=np.float32)
    b = np.ones(n, dtype=np.float32)
    a[0] = 2.0 ** 24
    got = zo.dot(a, b)
    # lane0 = 2^24 (+1 absorbed x3), lanes 1..15 = 4 each -> reduce: exact small sums then 2^24 + 60
    assert got == float(pyref.dot16(a, b))
    assert got == 2.0 ** 24 + 60.0
    seq = np.float32(0)
    for i in range(n):
        seq = np.float32(seq + a[i] * b[i])
    assert float(seq) != got  # the order is observable


def test_scalar_and_avx512_paths_agree_bitwise():
    rng = np.random.default_rng(7)
    rows = rng.standard_normal((64, 768)).astype(np.float32)
    qs = rng.standard_normal((64, 768)).astype(np.float32)
    res = {}
    for force in (1, 0):
        zo.lib().zbo_force_scalar(force)
        res[force] = [zo.distance_bits_batch(m, rows, qs) for m in (zo.COSINE, zo.L2SQ, zo.L2)]
        res[force].append(np.array([zo.dot(rows[i], qs[i]) for i in range(64)]))
    zo.lib().zbo_force_scalar(0)
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)


def test_integer_vectors_exact():
    a = np.arange(384, dtype=np.float32) % 7 - 3
    b = np.arange(384, dtype=np.float32) % 5 - 2
    assert zo.dot(a, b) == float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    assert zo.l2sq(a, b) == float(np.sum((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    assert zo.distance_bits(zo.L2SQ, a, b) == f64bits(np.sum((a.astype(np.float64) - b) ** 2))
    assert zo.distance_bits(zo.L2, a, b) == f64bits(np.sqrt(np.sum((a.astype(np.float64) - b) ** 2)))


def test_metric_bits_are_f64_bits_of_f32_sum():
    rng = np.random.default_rng(3)
    a = rng.standard_normal(768).astype(np.float32)
    b = rng.standard_normal(768).astype(np.float32)
    s = pyref.l2sq16(a, b)
    assert zo.distance_bits(zo.L2SQ, a, b) == f64bits(float(s))
    assert zo.distance_bits(zo.L2, a, b) == f64bits(np.sqrt(np.float64(s)))
    assert zo.distance_bits(zo.COSINE, a, b) == pyref.cos_zebra_bits(a, b)


def test_cosine_zebra_literal_semantics_q4():
    # distance.rs:23-25: (1.0 - simsimd cosine DISTANCE).to_bits()  -> the clipped similarity
    e0 = np.zeros(16, dtype=np.float32); e0[0] = 1
    e1 = np.zeros(16, dtype=np.float32); e1[1] = 1
    z = np.zeros(16, dtype=np.float32)
    assert zo.distance_bits(zo.COSINE, e0, e0) == f64bits(1.0)       # identical: distance 0 -> 1 - 0
    assert zo.distance_bits(zo.COSINE, e0, e1) == f64bits(0.0)       # orthogonal: ab == 0 -> distance 1 -> 0
    assert zo.distance_bits(zo.COSINE, e0, -e0) == f64bits(-1.0)     # opposite: distance 2 -> -1 (sign bit set)
    assert zo.distance_bits(zo.COSINE, z, z) == f64bits(1.0)         # both zero: distance 0 -> 1
    assert zo.distance_bits(zo.COSINE, z, e0) == f64bits(0.0)        # one zero: ab == 0 -> distance 1 -> 0
    # u64 ordering (lsh.rs:318): negative similarities carry the sign bit and sort last
    assert f64bits(0.0) < f64bits(1.0) < f64bits(-1.0)


def test_point_is_above_edges():
    n = 16
    zero = np.zeros(n, dtype=np.float32)
    x = np.ones(n, dtype=np.float32)
    assert zo.point_is_above(zero, -0.0, x) is True          # +0.0 + -0.0 = +0.0 >= 0
    assert zo.point_is_above(zero, 0.0, x) is True
    tiny = np.float32(1e-45)                                  # smallest subnormal
    assert zo.point_is_above(zero, -float(tiny), x) is False
    assert zo.point_is_above(zero, float(tiny), x) is True
    nan = zero.copy(); nan[3] = np.nan
    assert zo.point_is_above(nan, 0.0, x) is False            # NaN >= 0 is false
    c = zero.copy(); c[0] = 1.0
    p = zero.copy(); p[0] = -1.0
    assert zo.point_is_above(c, 1.0, p) is True               # -1 + 1 = 0 >= 0
    assert zo.point_is_above(c, np.nextafter(np.float32(1.0), np.float32(0.0)), p) is False


def test_make_plane_matches_reference_formula():
    rng = np.random.default_rng(11)
    a = rng.standard_normal(384).astype(np.float32)
    b = rng.standard_normal(384).astype(np.float32)
    coef, cst = zo.make_plane(a, b)
    exp_coef = (b - a).astype(np.float32)                      # lsh.rs:174-181, :222
    mid = ((a + b) / np.float32(2.0)).astype(np.float32)       # lsh.rs:183-190, :223
    assert np.array_equal(coef, exp_coef)
    assert np.float32(cst) == np.float32(-np.float32(pyref.dot16(exp_coef, mid)))  # lsh.rs:224-225
    assert zo.point_is_above(coef, cst, a) is False
    assert zo.point_is_above(coef, cst, b) is True
