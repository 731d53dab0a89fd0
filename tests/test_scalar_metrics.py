"""The ten scalar metrics of /root/reference/src/distance.rs:51-190, CPU side:

* the C oracle against an independent numpy restatement (strictly sequential f32 folds via ``np.add.accumulate``) and
  hand-derived known answers;
* the HOST TWIN of the device arithmetic (tests/metric_twin.cpp compiles zebra_b200/csrc/zb_metrics.cuh -- the very
  header the CUDA kernels include -- for the CPU) bit for bit against the oracle, adversarial inputs included;
* the deterministic p-th root both sides use in place of cbrtf / powf(., 1/p).

The GPU leg of the same comparison is tests/test_gpu_metrics.py.  PARITY UNPINNED: the `distances` crate is not on this
machine; its folds are restated from the published algorithm (oracle/README.md, "scalar metrics").
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import zb_oracle as zo

HERE = os.path.dirname(os.path.abspath(__file__))
F32 = np.float32

SCALAR = [
    ("chebyshev", zo.CHEBYSHEV), ("canberra", zo.CANBERRA), ("bray_curtis", zo.BRAY_CURTIS), ("manhattan", zo.MANHATTAN),
    ("l3", zo.L3), ("l4", zo.L4), ("hamming", zo.HAMMING), ("minkowski0", zo.MINKOWSKI(0)), ("minkowski1", zo.MINKOWSKI(1)),
    ("minkowski2", zo.MINKOWSKI(2)), ("minkowski3", zo.MINKOWSKI(3)), ("minkowski7", zo.MINKOWSKI(7)),
    ("minkowski64", zo.MINKOWSKI(64)), ("pnorm0", zo.PNORM(0)), ("pnorm2", zo.PNORM(2)), ("pnorm5", zo.PNORM(5)),
]


def f32bits(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def key(x) -> int:
    x = F32(x)
    return 0xFFC00000 if np.isnan(x) else f32bits(x)


def seqsum(terms) -> np.float32:
    terms = np.asarray(terms, dtype=F32)
    return np.add.accumulate(terms, dtype=F32)[-1] if terms.size else F32(0)


def powi(v, p):
    """compiler-rt __powisf2 on f32 arrays: square and multiply."""
    r = np.ones_like(v)
    a = v.copy()
    b = p
    while True:
        if b & 1:
            r = (r * a).astype(F32)
        b //= 2
        if b == 0:
            break
        a = (a * a).astype(F32)
    return r


def numpy_metric(code: int, a, b) -> int:
    """Independent restatement; roots taken from the oracle's root_p (checked separately below)."""
    a, b = np.asarray(a, F32), np.asarray(b, F32)
    c, power = code & 0xFF, code >> 8
    with np.errstate(all="ignore"):
        v = np.abs(a - b)
        if c == zo.CHEBYSHEV:
            acc = F32(0)
            for x in v:               # if acc > v { acc } else { v }
                acc = acc if acc > x else x
            return key(acc)
        if c == zo.CANBERRA:
            return key(seqsum(v / (np.abs(a) + np.abs(b))))
        if c == zo.BRAY_CURTIS:
            return key(seqsum(v) / seqsum(np.abs(a + b)))
        if c == zo.MANHATTAN:
            return key(seqsum(v))
        if c == zo.L3:
            return key(zo.root_p(seqsum((v * v) * v), 3))
        if c == zo.L4:
            v2 = v * v
            return key(np.sqrt(np.sqrt(seqsum(v2 * v2))))
        if c == zo.HAMMING:
            x = (a.view(np.uint32) ^ b.view(np.uint32)) & 0xFF
            return int(sum(bin(int(t)).count("1") for t in x))
        s = seqsum(powi(v, power))
        if c == 11:
            return key(s)
        if power == 0:
            return key(F32(1) if s == 1 else F32(np.inf))
        return key(zo.root_p(s, power))


def adversarial(rng, n, dim):
    a = rng.standard_normal((n, dim)).astype(F32)
    b = rng.standard_normal((n, dim)).astype(F32)
    a[0] = 0; b[0] = 0                       # both zero: Canberra / Bray-Curtis 0/0 -> NaN
    a[1] = 0                                 # one zero
    b[2] = a[2]                              # identical
    b[3] = -a[3]                             # opposite: Bray-Curtis denominator 0, numerator > 0 -> inf
    a[4] *= F32(1e-20); b[4] *= F32(1e-20)   # underflowing powers
    a[5] *= F32(1e18); b[5] *= F32(1e18)     # overflowing powers -> inf
    if dim > 2:
        a[6, 1] = 0; b[6, 1] = 0             # a single 0/0 term inside an otherwise ordinary row
        a[7, dim // 2] = np.nan              # NaN element (Chebyshev keeps it only if it is the last element)
        a[8, dim - 1] = np.nan
        a[9, 0] = np.inf
        a[10] = F32(1e-42); b[10] = F32(-1e-42)  # subnormals
    return a, b


# ------------------------------------------------------------------------------------------ oracle vs numpy restatement
@pytest.mark.parametrize("dim", [1, 3, 4, 5, 16, 20, 384])
def test_oracle_scalar_metrics_equal_numpy_restatement(dim):
    rng = np.random.default_rng(dim)
    a, b = adversarial(rng, 24, dim)
    for name, code in SCALAR:
        got = zo.distance_bits_batch(code, a, b)
        exp = np.array([numpy_metric(code, a[i], b[i]) for i in range(a.shape[0])], dtype=np.uint64)
        assert np.array_equal(got, exp), (name, dim, np.nonzero(got != exp)[0][:5])


def test_known_answers_by_hand():
    a = np.array([1, 2, 3, -4], F32)
    b = np.array([4, 0, 3, 4], F32)          # |a-b| = 3 2 0 8
    kb = lambda m: zo.distance_bits(m, a, b)
    assert kb(zo.MANHATTAN) == f32bits(13.0)
    assert kb(zo.CHEBYSHEV) == f32bits(8.0)
    # 3/5 + 2/2 + 0/6 + 8/8, folded left to right in f32
    assert kb(zo.CANBERRA) == f32bits(F32(F32(F32(F32(3) / F32(5)) + F32(1)) + F32(0)) + F32(1))
    assert kb(zo.BRAY_CURTIS) == f32bits(F32(13) / F32(13))          # |a+b| = 5 2 6 0
    assert kb(zo.L3) == f32bits(zo.root_p(27 + 8 + 512, 3))
    assert kb(zo.L4) == f32bits(np.sqrt(np.sqrt(F32(81 + 16 + 4096))))
    assert kb(zo.PNORM(2)) == f32bits(9 + 4 + 64) and kb(zo.MINKOWSKI(2)) == f32bits(zo.root_p(77.0, 2))
    assert kb(zo.PNORM(0)) == f32bits(4.0)                           # powi(v, 0) = 1 for every element, 0 included
    assert kb(zo.MINKOWSKI(0)) == 0x7F800000                         # powf(4, 1/0) = +inf (MinkowskiDistance::default())
    assert kb(zo.MINKOWSKI(1)) == kb(zo.MANHATTAN)
    # Hamming looks at the LOW BYTE of each f32 bit pattern only (distance.rs:147-148): small integers share low byte 0
    assert kb(zo.HAMMING) == 0
    x = np.array([0.1, 1.0], F32)            # 0x3DCCCCCD, 0x3F800000
    y = np.array([0.2, 1.5], F32)            # 0x3E4CCCCD, 0x3FC00000
    assert zo.distance_bits(zo.HAMMING, x, y) == 0
    y = np.array([0.3, 1.0], F32)            # 0x3E99999A: low bytes CD ^ 9A = 57 -> 5 bits
    assert zo.distance_bits(zo.HAMMING, x, y) == 5
    # both elements zero: Canberra's 0/0 poisons the sum with the x86 default NaN; order by bits puts it after +inf
    z = np.zeros(3, F32)
    assert zo.distance_bits(zo.CANBERRA, z, z) == 0xFFC00000 and zo.distance_bits(zo.BRAY_CURTIS, z, z) == 0xFFC00000
    assert zo.distance_bits(zo.MANHATTAN, z, z) == 0 and zo.distance_bits(zo.CHEBYSHEV, z, z) == 0


def test_sequential_order_is_observable():
    """The fold order is part of the contract: a sum that depends on it must come out in input order."""
    a = np.array([16777216.0, 1.0, 1.0], F32)     # 2^24: adding 1 is a tie that rounds back to even
    b = np.zeros(3, F32)
    first = zo.distance_bits(zo.MANHATTAN, a, b)            # (2^24 + 1) + 1 = 2^24
    last = zo.distance_bits(zo.MANHATTAN, a[[1, 2, 0]], b)  # (1 + 1) + 2^24 = 2^24 + 2
    assert first == f32bits(16777216.0) and last == f32bits(16777218.0)


# ------------------------------------------------------------------------------------------ the deterministic root
def test_root_p_exact_on_perfect_powers_and_within_one_ulp_of_libm():
    for p in (2, 3, 4, 5, 7, 10, 31, 64):
        for r in (1.0, 2.0, 3.0, 0.5, 10.0):
            with np.errstate(over="ignore"):
                s = F32(r) ** p
            if np.isfinite(s) and F32(float(s)) == s and float(s) == float(r) ** p:
                assert zo.root_p(float(s), p) == r, (p, r)
    rng = np.random.default_rng(3)
    s = np.exp(rng.uniform(-80, 80, 4000)).astype(F32)
    for p in (2, 3, 7, 64):
        got = np.array([zo.root_p(float(x), p) for x in s], dtype=F32)
        true = (s.astype(np.float64) ** (1.0 / p))
        ulp = np.spacing(true.astype(F32)).astype(np.float64)
        assert np.all(np.abs(got.astype(np.float64) - true) <= 0.5000001 * ulp), p       # correctly rounded on this sample
    got = np.array([zo.root_p(float(x), 2) for x in s], dtype=F32)
    assert np.array_equal(got, np.sqrt(s.astype(np.float64)).astype(F32))
    assert zo.root_p(0.0, 3) == 0.0 and zo.root_p(float("inf"), 3) == float("inf") and np.isnan(zo.root_p(float("nan"), 3))
    assert zo.root_p(1e-45, 3) > 0 and zo.root_p(3.4028235e38, 64) > 1.0
    xs = np.sort(s)
    for p in (3, 5):
        ys = np.array([zo.root_p(float(x), p) for x in xs], dtype=F32)
        assert np.all(np.diff(ys) >= 0)                                                  # monotone: order by bits is kept


# ------------------------------------------------------------------------------------------ host twin of the device code
@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("twin") / "libmetric_twin.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", out,
                           os.path.join(HERE, "metric_twin.cpp")])
    L = C.CDLL(out)
    L.twin_distance_bits_batch.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.twin_root_p.restype = C.c_float
    L.twin_root_p.argtypes = [C.c_float, C.c_int]
    return L


def twin_bits(L, code, a, b):
    a, b = np.ascontiguousarray(a, F32), np.ascontiguousarray(b, F32)
    out = np.empty(a.shape[0], np.uint64)
    assert L.twin_distance_bits_batch(code & 0xFF, code >> 8, a.shape[0], a.ctypes.data, b.ctypes.data, a.shape[1],
                                      out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 7, 16, 20, 384, 768])
def test_device_arithmetic_twin_equals_oracle(twin, dim):
    rng = np.random.default_rng(1000 + dim)
    a, b = adversarial(rng, 300, dim)
    a[20:60] = np.round(a[20:60] * 4) / 4           # coarse values: many exact ties and cancellations
    b[20:60] = np.round(b[20:60] * 4) / 4
    for name, code in SCALAR:
        assert np.array_equal(twin_bits(twin, code, a, b), zo.distance_bits_batch(code, a, b)), (name, dim)


def test_device_root_twin_equals_oracle(twin):
    rng = np.random.default_rng(9)
    s = np.concatenate([np.exp(rng.uniform(-100, 88, 20000)), [0.0, 1e-45, 1.0, 3.4028235e38, np.inf]]).astype(F32)
    for p in (1, 2, 3, 4, 5, 7, 13, 64):
        for x in s[:: (1 if p in (3, 7) else 20)]:
            assert f32bits(twin.twin_root_p(float(x), p)) == f32bits(zo.root_p(float(x), p)), (p, float(x))


# ------------------------------------------------------------------------------------------ the walk with a scalar metric
@pytest.mark.parametrize("name,code", [("manhattan", zo.MANHATTAN), ("chebyshev", zo.CHEBYSHEV), ("hamming", zo.HAMMING),
                                       ("minkowski3", zo.MINKOWSKI(3))])
def test_oracle_search_with_scalar_metric_is_sorted_and_matches_brute_force_over_candidates(name, code):
    rng = np.random.default_rng(7)
    rows = rng.standard_normal((600, 24)).astype(F32)
    orc = zo.OracleIndex(24, code, 8, 4, seed=5)
    orc.add(rows)
    for q in rows[:10]:
        ids, bits = orc.search(q, 10)
        cand = orc.candidates(q, 10)
        exp = sorted((zo.distance_bits(code, rows[i], q), int(i)) for i in cand)[:10]
        assert [(int(b), int(i)) for i, b in zip(ids, bits)] == exp
