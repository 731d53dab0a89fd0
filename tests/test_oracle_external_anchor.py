"""External anchor for the oracle's METRIC VALUES: oracle/zb_oracle.c restates the arithmetic of the
un-vendored crates simsimd / distances from their published algorithms, and nothing in /root/reference holds a golden
value for them.  Here every metric of /root/reference/src/distance.rs:19-190 is compared with two implementations that
share no code with the oracle: numpy in f64 (written from the textbook definitions below) and scipy.spatial.distance.
Bar = north_star's: 1e-5 relative (+1e-7 absolute near 0); the f32 accumulation of the reference's kernels is itself only
that close to the exact value.  Random, clustered and adversarial vectors (tiny, huge, identical, opposite, zero).

This anchors the VALUE each metric returns; it cannot anchor the last-bit behaviour of the reference's SIMD kernels (that
would need the reference built here: no cargo / rustc in this image, see test_reference_toolchain_probe)."""
import os
import shutil
import struct

import numpy as np
import pytest
from scipy.spatial import distance as sd

from oracle import zb_oracle as zo

F32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits_to_f64(b):
    return struct.unpack("<d", struct.pack("<Q", int(b)))[0]


def bits_to_f32(b):
    return struct.unpack("<f", struct.pack("<I", int(b) & 0xFFFFFFFF))[0]


def vectors(dim, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((96, dim)).astype(F32)
    b = rng.standard_normal((96, dim)).astype(F32)
    c = rng.standard_normal((1, dim)).astype(F32)
    a[10:20] = c + 0.05 * a[10:20]            # near neighbours (clustered data: small distances, cosine near 1)
    b[10:20] = c + 0.05 * b[10:20]
    a[20] = b[20]                             # identical
    a[21] = -b[21]                            # opposite
    a[22] *= F32(1e-6); b[22] *= F32(1e-6)    # small
    a[23] *= F32(1e6); b[23] *= F32(1e6)      # large
    a[24] = np.abs(a[24]); b[24] = np.abs(b[24])
    return a, b


def close(got, exp, rel=1e-5, abs_=1e-7):
    if np.isinf(exp) or np.isnan(exp):      # e.g. Bray-Curtis of opposite vectors: |a - b| / |a + b| = x / 0
        return (np.isinf(got) and got == exp) or (np.isnan(got) and np.isnan(exp))
    return abs(got - exp) <= rel * abs(exp) + abs_


@pytest.mark.parametrize("dim", [16, 100, 384, 768])
def test_simsimd_path_metrics_against_numpy_f64_and_scipy(dim):
    a, b = vectors(dim, dim)
    for i in range(a.shape[0]):
        x, y = a[i].astype(np.float64), b[i].astype(np.float64)
        l2sq = float(np.sum((x - y) ** 2))
        # distance.rs:38-49 / :103-114: simsimd l2sq / l2, f64 bits
        got = bits_to_f64(zo.distance_bits(zo.L2SQ, a[i], b[i]))
        assert close(got, l2sq) and close(got, float(sd.sqeuclidean(x, y))), (i, got, l2sq)
        got = bits_to_f64(zo.distance_bits(zo.L2, a[i], b[i]))
        assert close(got, np.sqrt(l2sq)) and close(got, float(sd.euclidean(x, y))), (i, got)
        # distance.rs:19-32 (quirk Q4): 1.0 - (simsimd cosine DISTANCE clipped at 0) = min(similarity, 1)
        sim = float(np.dot(x, y) / (np.linalg.norm(x) * np.linalg.norm(y)))
        got = bits_to_f64(zo.distance_bits(zo.COSINE, a[i], b[i]))
        assert close(got, min(sim, 1.0), abs_=2e-6), (i, got, sim)
        assert close(got, min(1.0 - float(sd.cosine(x, y)), 1.0), abs_=2e-6), (i, got)
        assert got <= 1.0


@pytest.mark.parametrize("dim", [16, 100, 384])
def test_distances_crate_metrics_against_numpy_f64_and_scipy(dim):
    a, b = vectors(dim, 1000 + dim)
    for i in range(a.shape[0]):
        x, y = a[i].astype(np.float64), b[i].astype(np.float64)
        v = np.abs(x - y)
        exp = {
            zo.CHEBYSHEV: (float(v.max()), float(sd.chebyshev(x, y))),
            zo.MANHATTAN: (float(v.sum()), float(sd.cityblock(x, y))),
            zo.L3: (float(np.sum(v ** 3) ** (1.0 / 3.0)), float(sd.minkowski(x, y, 3))),
            zo.L4: (float(np.sum(v ** 4) ** 0.25), float(sd.minkowski(x, y, 4))),
            zo.BRAY_CURTIS: (float(v.sum() / np.abs(x + y).sum()), float(sd.braycurtis(x, y))),
            zo.MINKOWSKI(3): (float(np.sum(v ** 3) ** (1.0 / 3.0)), float(sd.minkowski(x, y, 3))),
            zo.PNORM(3): (float(np.sum(v ** 3)), float(sd.minkowski(x, y, 3)) ** 3),
        }
        if not np.any((np.abs(x) + np.abs(y)) == 0):
            exp[zo.CANBERRA] = (float(np.sum(v / (np.abs(x) + np.abs(y)))), float(sd.canberra(x, y)))
        for mid, (e_np, e_sp) in exp.items():
            got = bits_to_f32(zo.distance_bits(mid, a[i], b[i]))     # quirk Q6: f32 bits zero-extended
            assert close(got, e_np, rel=2e-5) and close(got, e_sp, rel=2e-5), (mid, i, got, e_np, e_sp)
        # distance.rs:147-155: Hamming counts the differing bits of the LOW BYTE of each f32's bit pattern; the integer
        # count itself is the DistanceUnit (`.into()`), not a float's bits
        lo = (a[i].view(np.uint32) ^ b[i].view(np.uint32)) & 0xFF
        assert zo.distance_bits(zo.HAMMING, a[i], b[i]) == sum(bin(int(w)).count("1") for w in lo)


def test_point_is_above_against_f64(dim=768):
    """lsh.rs:39-43 against an f64 dot product: the sign may only differ where the f64 value is within f32 rounding of 0."""
    rng = np.random.default_rng(5)
    coef = rng.standard_normal((2000, dim)).astype(F32)
    x = rng.standard_normal((2000, dim)).astype(F32)
    cst = rng.standard_normal(2000).astype(F32)
    got = zo.above_batch(coef, cst, x).astype(bool)
    exact = np.einsum("ij,ij->i", coef.astype(np.float64), x.astype(np.float64)) + cst.astype(np.float64)
    scale = np.einsum("ij,ij->i", np.abs(coef).astype(np.float64), np.abs(x).astype(np.float64))
    sure = np.abs(exact) > 1e-5 * scale
    assert sure.mean() > 0.99
    assert np.array_equal(got[sure], exact[sure] >= 0)


def test_reference_toolchain_probe():
    """use the reference itself as the oracle's pin the moment it can be built or found.  Today neither a Rust
    toolchain nor a driver-installed baseline/_ref exists, so parity stays UNPINNED and oracle/README.md must say so."""
    have_cargo = shutil.which("cargo") is not None and shutil.which("rustc") is not None
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref")) or os.path.isdir(os.path.join(ROOT, "oracle", "_ref"))
    readme = open(os.path.join(ROOT, "oracle", "README.md")).read()
    if not have_cargo and not have_ref:
        assert "PARITY UNPINNED" in readme
        pytest.skip("no cargo/rustc and no baseline/_ref: the reference cannot be run here; parity unpinned (anchored on numpy f64 / scipy above)")
    pytest.fail("a Rust toolchain or baseline/_ref appeared: build the reference under oracle/_ref and pin the oracle against it")
