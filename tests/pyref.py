"""Independent pure-Python restatements used to pin the C oracle (tests only, small cases).

* exact-rational emulation of f32 fused multiply-add and of the "skylake-16" accumulation order
  (simsimd AVX-512 f32 kernels as called from /root/reference/src/database/index/lsh.rs:40 and
  /root/reference/src/distance.rs:23,41,106);
* a second, recursion-for-recursion restatement of tree_result / search
  (/root/reference/src/database/index/lsh.rs:290-348, :544-565) over the flat forest arrays.
"""
from __future__ import annotations

import struct
from fractions import Fraction

import numpy as np


def _f32_to_frac(x) -> Fraction:
    return Fraction(float(np.float32(x)))


def round_f32(fr: Fraction) -> np.float32:
    """Round an exact rational to the nearest f32, ties to even (subnormals handled, no overflow handling)."""
    if fr == 0:
        return np.float32(0.0)
    sign = -1 if fr < 0 else 1
    a = abs(fr)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    elif Fraction(2) ** (e + 1) <= a:
        e += 1
    e = max(e, -126)
    scale = Fraction(2) ** (e - 23)
    m = a / scale
    fl = m.numerator // m.denominator
    rem = m - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (fl & 1)):
        fl += 1
    return np.float32(sign * float(Fraction(fl) * scale))


def fma32(a, b, c) -> np.float32:
    return round_f32(_f32_to_frac(a) * _f32_to_frac(b) + _f32_to_frac(c))


def add32(a, b) -> np.float32:
    return np.float32(np.float32(a) + np.float32(b))


def reduce16(acc):
    x = [add32(acc[i], acc[i + 8]) for i in range(8)]
    r = [add32(x[i], x[i + 4]) for i in range(4)]
    return add32(add32(r[0], r[1]), add32(r[2], r[3]))


def dot16(a, b) -> np.float32:
    acc = [np.float32(0.0)] * 16
    for i in range(len(a)):
        acc[i % 16] = fma32(a[i], b[i], acc[i % 16])
    return reduce16(acc)


def l2sq16(a, b) -> np.float32:
    acc = [np.float32(0.0)] * 16
    for i in range(len(a)):
        d = np.float32(np.float32(a[i]) - np.float32(b[i]))
        acc[i % 16] = fma32(d, d, acc[i % 16])
    return reduce16(acc)


def bits64(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def cos_zebra_bits(a, b) -> int:
    ab, a2, b2 = float(dot16(a, b)), float(dot16(a, a)), float(dot16(b, b))
    if a2 == 0.0 and b2 == 0.0:
        c = 0.0
    elif ab == 0.0:
        c = 1.0
    else:
        t = (ab * (1.0 / np.sqrt(np.float64(a2)))) * (1.0 / np.sqrt(np.float64(b2)))
        c = max(0.0, 1.0 - float(t))
    return bits64(1.0 - c)


class PyForest:
    """tree_result / search restated in Python over flat forest arrays; arithmetic via callables."""

    def __init__(self, forest, rows, tomb, dist_bits, is_above):
        self.f, self.rows, self.tomb = forest, rows, tomb
        self.dist_bits, self.is_above = dist_bits, is_above

    def members(self, leaf):
        lo, hi = int(self.f.leaf_off[leaf]), int(self.f.leaf_off[leaf + 1])
        return [int(i) for i in self.f.members[lo:hi] if not self.tomb[int(i)]]

    def tree_result(self, q, n, node, cand, trace, tree):
        plane, left, right, leaf = (int(v) for v in self.f.nodes[node])
        if plane < 0:
            mem = self.members(leaf)
            trace.append((tree, leaf, n, len(mem)))
            if len(mem) < n:
                cand.update(mem)
                return len(mem)
            scored = sorted((self.dist_bits(self.rows[i], q), i) for i in mem)
            cand.update(i for _, i in scored[:n])
            return n
        above = self.is_above(self.f.coef[plane], float(self.f.cst[plane]), q)
        main, backup = (right, left) if above else (left, right)
        k = self.tree_result(q, n, main, cand, trace, tree)
        if k < n:
            return self.tree_result(q, n - k, backup, cand, trace, tree)
        return k

    def search(self, q, top_k):
        cand, trace = set(), []
        for t, root in enumerate(self.f.roots):
            self.tree_result(q, top_k, int(root), cand, trace, t)
        scored = sorted((self.dist_bits(self.rows[i], q), i) for i in cand)
        return scored[:top_k], sorted(cand), trace
