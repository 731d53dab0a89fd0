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
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", out,
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


# ------------------------------------------------------------------------------------------ the kernel source, emulated
@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    """tests/quadtile_emu.cpp: quad_tile_kernel's own source compiled for the CPU, one std::thread per CUDA thread."""
    out = str(tmp_path_factory.mktemp("qtemu") / "libquadtile_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread",
                           "-fvisibility=hidden", "-Wl,-Bsymbolic", "-o", out, os.path.join(HERE, "quadtile_emu.cpp")])
    L = C.CDLL(os.environ.get("ZB_QUADTILE_EMU_SO", out))     # a prebuilt (sanitizer) build may be substituted
    L.emu_quad_tile.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_uint32] + [C.c_void_p] * 9
    return L


@pytest.mark.parametrize("metric,dim", [(0, 20), (1, 48), (2, 100), (0, 16)])
def test_quad_tile_kernel_source_emulated_on_cpu(emu, metric, dim):
    """Whole tiles through the kernel's source: leaves of 1..300 rows (tails of every size), 0..19 visits per leaf (tiles
    of 1..8 queries, one or two query groups), tombstones, dim % 16 != 0 -- every key against the oracle."""
    rng = np.random.default_rng(metric * 100 + dim)
    dimp = (dim + 15) // 16 * 16
    leaf_sizes = [1, 3, 4, 5, 64, 129, 300, 7, 256]
    n = sum(leaf_sizes)
    rows = rng.standard_normal((n, dim)).astype(F32)
    rows[10] = 0                                               # a zero row (cosine special case)
    rows_p = np.zeros((n, dimp), F32); rows_p[:, :dim] = rows
    slots = rng.permutation(n).astype(np.uint32)               # members are slots in any order
    leaf_off = np.concatenate([[0], np.cumsum(leaf_sizes)]).astype(np.int64)
    leaf_len = np.array(leaf_sizes, np.uint32)
    tomb_bits = rng.random(n) < 0.15
    tomb = np.zeros(n // 32 + 1, np.uint32)
    for s in np.nonzero(tomb_bits)[0]:
        tomb[s >> 5] |= np.uint32(1 << (s & 31))
    nq = 40
    queries = rng.standard_normal((nq, dim)).astype(F32)
    queries[0] = 0
    queries[1] = rows[int(slots[0])]                           # an identical pair
    q_p = np.zeros((nq, dimp), F32); q_p[:, :dim] = queries
    # visits in plan order (shuffled over leaves), grouped by leaf into tiles of <= 8 as the device grouping does
    per_leaf = [int(rng.integers(0, 20)) for _ in leaf_sizes]
    per_leaf[6] = 17                                           # the 300-row leaf: two full tiles + one of 1
    v_leaf = np.repeat(np.arange(len(leaf_sizes)), per_leaf)
    rng.shuffle(v_leaf)
    nv = v_leaf.size
    v_q = rng.integers(0, nq, nv).astype(np.uint32)
    pair_off = np.concatenate([[0], np.cumsum(leaf_len[v_leaf])]).astype(np.uint64)
    order, tile_leaf, tile_first, tile_count = [], [], [], []
    for l in range(len(leaf_sizes)):
        vs = [v for v in range(nv) if v_leaf[v] == l]
        rng.shuffle(vs)
        for i in range(0, len(vs), 8):
            tile_leaf.append(l); tile_first.append(len(order) + i); tile_count.append(min(8, len(vs) - i))
        order += vs
    order = np.array(order or [0], np.uint32)
    tl, tf, tc = (np.array(x or [0], np.uint32) for x in (tile_leaf, tile_first, tile_count))
    pair_key = np.full(int(pair_off[-1]) + 1, 0x1234, np.uint64)
    stats = np.zeros(3, np.uint64)
    done = emu.emu_quad_tile(metric, 2, dim, leaf_off.ctypes.data, leaf_len.ctypes.data, slots.ctypes.data, rows_p.ctypes.data,
                             tomb.ctypes.data, len(tile_leaf), tl.ctypes.data, tf.ctypes.data, tc.ctypes.data, order.ctypes.data,
                             v_q.ctypes.data, pair_off.ctypes.data, q_p.ctypes.data, pair_key.ctypes.data, stats.ctypes.data)
    assert done >= len(tile_leaf)                               # every tile was taken (each block's last fetch overshoots)
    exp = np.full(pair_key.size, 0x1234, np.uint64)
    for v in range(nv):
        l = int(v_leaf[v])
        for r in range(leaf_sizes[l]):
            slot = int(slots[int(leaf_off[l]) + r])
            exp[int(pair_off[v]) + r] = (0xFFFFFFFFFFFFFFFF if tomb_bits[slot]
                                         else zo.distance_bits(metric, rows[slot], queries[int(v_q[v])]))
    bad = np.nonzero(pair_key != exp)[0]
    assert bad.size == 0, (bad[:10], pair_key[bad[:3]], exp[bad[:3]])
    assert int(stats[0]) == nv and int(stats[1]) == int(pair_off[-1])
    assert int(stats[2]) == sum((leaf_sizes[l] + c) * 4 * dim for l, c in zip(tile_leaf, tile_count))
