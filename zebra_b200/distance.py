"""Distance metrics: the host mirror of /root/reference/src/distance.rs.

The three metrics on the north-star path keep the reference's names and its trait shape --
``metric.distance(a, b) -> DistanceUnit`` (u64 = f64 bit pattern, distance.rs:13) -- and additionally carry
the ``METRIC`` enum the device index needs.  They are evaluated by the CUDA library (no CPU arithmetic here).
The reference's ten other metrics (distance.rs:51-190) are outside the hot-path scope.
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

    def distance(self, a, b) -> DistanceUnit:
        """Metric::distance(&self, a: &Embedding<N>, b: &Embedding<N>) -> u64, arguments (stored row, query)."""
        return int(self.distance_batch(np.asarray(a, np.float32)[None, :], np.asarray(b, np.float32)[None, :])[0])

    def distance_batch(self, a, b) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        if a.shape != b.shape or a.ndim != 2:
            raise ValueError("distance_batch takes two [n, dim] arrays of the same shape")
        out = np.empty(a.shape[0], dtype=np.uint64)
        _ffi.check(_ffi.lib().zb_metric_distance_batch(self.device, self.METRIC, a.shape[0], a.shape[1],
                                                       a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)


class CosineDistance(_DeviceMetric):
    """distance.rs:15-32.  Literal semantics: (1.0 - simsimd cosine distance).to_bits() (survey quirk Q4)."""
    METRIC = _ffi.METRIC_COSINE


class L2SquaredDistance(_DeviceMetric):
    """distance.rs:34-49."""
    METRIC = _ffi.METRIC_L2SQ


class L2Distance(_DeviceMetric):
    """distance.rs:99-114."""
    METRIC = _ffi.METRIC_L2


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
