"""Host twin of the keys-only leaf-tile scan's register tile (zebra_b200/csrc/zb_quadtile.cuh, quad_tile_kernel: cosine /
L2 visits with n' > 32, BASELINE config 5's top-100): tests/quadtile_twin.cpp replays a quad on the CPU -- qt_chunk per
thread, then the quad_reduce16 fold -- and the sums must be the oracle's, bit for bit (Metric::distance of
/root/reference/src/distance.rs:19-49, :103-114 in the canonical order)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import zb_oracle as zo

HERE = os.path.dirname(os.path.abspath(__file__))
F32 = np.float32


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("qt") / "libquadtile_twin.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "quadtile_twin.cpp")])
    L = C.CDLL(out)
    L.twin_quadtile.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("n,m,dim", [(1, 1, 16), (5, 3, 20), (37, 8, 100), (64, 5, 384), (9, 8, 768)])
def test_quad_tile_twin_equals_oracle_sums(twin, n, m, dim):
    rng = np.random.default_rng(n * m + dim)
    rows = rng.standard_normal((n, dim)).astype(F32)
    queries = rng.standard_normal((m, dim)).astype(F32)
    rows[0] *= F32(1e-20)
    queries[-1] *= F32(1e19)
    if n > 2:
        rows[2] = queries[0]                       # identical pair: L2 sum exactly 0
    lib = zo.lib()
    lib.zbo_cos3_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    for metric in (0, 1):
        out_m = np.zeros((n, m), F32)
        a2 = np.zeros(n, F32)
        b2 = np.zeros(m, F32)
        twin.twin_quadtile(metric, rows.ctypes.data, n, queries.ctypes.data, m, dim, out_m.ctypes.data, a2.ctypes.data, b2.ctypes.data)
        exp = np.zeros((n, m), F32)
        ea2, eb2 = np.zeros(n, F32), np.zeros(m, F32)
        for i in range(n):
            for j in range(m):
                if metric == 0:
                    ab, x2, y2 = C.c_float(), C.c_float(), C.c_float()
                    lib.zbo_cos3_f32(rows[i].ctypes.data, queries[j].ctypes.data, dim, C.byref(ab), C.byref(x2), C.byref(y2))
                    exp[i, j], ea2[i], eb2[j] = ab.value, x2.value, y2.value
                else:
                    exp[i, j] = F32(zo.l2sq(rows[i], queries[j]))
        assert np.array_equal(out_m.view(np.uint32), exp.view(np.uint32)), metric
        if metric == 0:
            assert np.array_equal(a2.view(np.uint32), ea2.view(np.uint32)) and np.array_equal(b2.view(np.uint32), eb2.view(np.uint32))
