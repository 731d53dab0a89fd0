"""Distance metrics: the host mirror of /root/reference/src/distance.rs.

Every metric keeps the reference's name and its trait shape -- ``metric.distance(a, b) -> DistanceUnit`` (u64,
distance.rs:13) -- and additionally carries the ``METRIC`` enum the device index needs.  They are evaluated by the CUDA
library (no CPU arithmetic here).  Cosine / L2Squared / L2 (the north-star path, simsimd) return f64 bit patterns; the
ten scalar metrics (distance.rs:51-190, the `distances` crate) return f32 bit patterns zero-extended, exactly as the
reference's ``.to_bits().into()`` does (survey quirk Q6).
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _ffi

DistanceUnit = int  # u64


class _DeviceMetric:
    METRIC: int = -1
    device: int = 0
    power: int = 0          # MinkowskiDistance / PNormDistance only
    BITS: int = 64          # width of the IEEE pattern inside DistanceUnit

    def distance(self, a, b) -> DistanceUnit:
        """Metric::distance(&self, a: &Embedding<N>, b: &Embedding<N>) -> u64, arguments (stored row, query)."""
        return int(self.distance_batch(np.asarray(a, np.float32)[None, :], np.asarray(b, np.float32)[None, :])[0])

    def distance_batch(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        if a.shape != b.shape or a.ndim != 2:
            raise ValueError("distance_batch takes two [n, dim] arrays of the same shape")
        out = np.empty(a.shape[0], dtype=np.uint64)
        _ffi.check(_ffi.lib().zb_metric_distance_batch(self.device, self.METRIC, self.power, a.shape[0], a.shape[1],
                                                       a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def to_float(self, bits) -> np.ndarray:
        """DistanceUnit -> the floating-point value it encodes (Hamming: the count itself)."""
        bits = np.asarray(bits, dtype=np.uint64)
        if self.BITS == 64:
            return bits.view(np.float64)
        if self.BITS == 32:
            return bits.astype(np.uint32).view(np.float32)
        return bits.astype(np.float64)

    def __eq__(self, other):
        return type(self) is type(other) and self.power == other.power

    def __hash__(self):
        return hash((type(self).__name__, self.power))


class CosineDistance(_DeviceMetric):
    """distance.rs:15-32.  Literal semantics: (1.0 - simsimd cosine distance).to_bits() (survey quirk Q4)."""
    METRIC = _ffi.METRIC_COSINE


class L2SquaredDistance(_DeviceMetric):
    """distance.rs:34-49."""
    METRIC = _ffi.METRIC_L2SQ


class L2Distance(_DeviceMetric):
    """distance.rs:99-114."""
    METRIC = _ffi.METRIC_L2


class _ScalarMetric(_DeviceMetric):
    BITS = 32


class ChebyshevDistance(_ScalarMetric):
    """distance.rs:51-61."""
    METRIC = _ffi.METRIC_CHEBYSHEV


class CanberraDistance(_ScalarMetric):
    """distance.rs:63-73.  An element pair (0, 0) contributes 0/0 = NaN, as IEEE arithmetic gives the reference."""
    METRIC = _ffi.METRIC_CANBERRA


class BrayCurtisDistance(_ScalarMetric):
    """distance.rs:75-85."""
    METRIC = _ffi.METRIC_BRAY_CURTIS


class ManhattanDistance(_ScalarMetric):
    """distance.rs:87-97."""
    METRIC = _ffi.METRIC_MANHATTAN


class L3Distance(_ScalarMetric):
    """distance.rs:116-126."""
    METRIC = _ffi.METRIC_L3


class L4Distance(_ScalarMetric):
    """distance.rs:128-138."""
    METRIC = _ffi.METRIC_L4


class HammingDistance(_ScalarMetric):
    """distance.rs:140-157: popcount over the low byte of every element's f32 bit pattern (`x.to_bits() as u8`)."""
    METRIC = _ffi.METRIC_HAMMING
    BITS = 0


class _PowerMetric(_ScalarMetric):
    def __init__(self, power: int = 0):
        """`power` as the reference's public field (distance.rs:164, :181); `Default` gives 0."""
        power = int(power)
        if not 0 <= power <= _ffi.METRIC_MAX_POWER:
            raise ValueError(f"power must be in 0..{_ffi.METRIC_MAX_POWER}")
        self.power = power


class MinkowskiDistance(_PowerMetric):
    """distance.rs:159-173."""
    METRIC = _ffi.METRIC_MINKOWSKI


class PNormDistance(_PowerMetric):
    """distance.rs:175-190."""
    METRIC = _ffi.METRIC_PNORM


def bits_to_f64(bits) -> np.ndarray:
    return np.asarray(bits, dtype=np.uint64).view(np.float64)


def f64_to_bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def point_is_above(coef, constant, x, device: int = 0) -> np.ndarray:
    """Hyperplane::point_is_above (lsh.rs:39-43) for n (plane, point) pairs, on the device."""
    coef = np.ascontiguousarray(np.atleast_2d(coef), dtype=np.float32)
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    cst = np.ascontiguousarray(np.atleast_1d(constant), dtype=np.float32)
    out = np.empty(coef.shape[0], dtype=np.uint8)
    _ffi.check(_ffi.lib().zb_point_is_above_batch(device, coef.shape[0], coef.shape[1], coef.ctypes.data,
                                                  cst.ctypes.data, x.ctypes.data, out.ctypes.data))
    return out.astype(bool)
