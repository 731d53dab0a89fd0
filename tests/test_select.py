"""warp_select_visits_kernel (zebra_b200/csrc/zb_select_kernel.cuh; knob select_variant = 1): the per-visit top-n' of
/root/reference/src/database/index/lsh.rs:301-331 with the list in registers, one warp per visit.  tests/select_emu.cpp runs
the kernel's OWN SOURCE on the CPU (one std::thread per lane, shuffles and ballots through a mailbox); every visit's output
must be the first n' entries of the visit's (key, ordinal)-sorted live members, tombstones (sentinel keys) left out, unused
slots filled with the sentinel.  The GPU leg is in tests/test_zz_quad_tile_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SENT = 0xFFFFFFFFFFFFFFFF


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("selemu") / "libselect_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-Wl,-Bsymbolic", "-o", out,
                           os.path.join(HERE, "select_emu.cpp")])
    L = C.CDLL(os.environ.get("ZB_SELECT_EMU_SO", out))     # a prebuilt (sanitizer) build may be substituted
    L.emu_warp_select.argtypes = [C.c_int, C.c_uint32] + [C.c_void_p] * 11
    return L


@pytest.mark.parametrize("KL,kmax", [(1, 10), (1, 32), (2, 64), (4, 100), (4, 128)])
def test_warp_select_kernel_source_emulated_on_cpu(emu, KL, kmax):
    rng = np.random.default_rng(KL * 1000 + kmax)
    leaf_sizes = [0, 1, 5, 31, 32, 33, 200, 700, 64]
    n = sum(leaf_sizes)
    leaf_off = np.concatenate([[0], np.cumsum(leaf_sizes)]).astype(np.int64)
    leaf_len = np.array(leaf_sizes, np.uint32)
    members = rng.permutation(n).astype(np.uint32)                 # slots
    ordinals = rng.permutation(10 * n)[:n].astype(np.uint64)       # ord[slot]: arbitrary distinct ordinals
    visits = []                                                    # (leaf, n', key generator)
    for l in range(len(leaf_sizes)):
        for mode in ("random", "ties", "descending", "mostly_dead", "all_dead"):
            visits.append((l, int(rng.integers(1, kmax + 1)), mode))
    visits.append((7, kmax, "random"))
    visits.append((6, 0, "random"))                                # n' = 0: nothing to keep
    nv = len(visits)
    vleaf = np.array([v[0] for v in visits], np.uint32)
    vnp = np.array([v[1] for v in visits], np.uint32)
    pair_off = np.concatenate([[0], np.cumsum(leaf_len[vleaf])]).astype(np.uint64)
    pair_key = np.zeros(int(pair_off[-1]) + 1, np.uint64)
    live_cnt = []
    for v, (l, k, mode) in enumerate(visits):
        m = leaf_sizes[l]
        if mode == "random":
            keys = rng.integers(0, 1 << 62, m).astype(np.uint64)
        elif mode == "ties":
            keys = rng.integers(0, 4, m).astype(np.uint64)         # many equal keys: order by ordinal
        elif mode == "descending":
            keys = np.arange(m, 0, -1).astype(np.uint64)           # every key beats the list: worst case for insertion
        else:
            keys = rng.integers(0, 1 << 40, m).astype(np.uint64)
        dead = np.zeros(m, bool)
        if mode == "mostly_dead":
            dead = rng.random(m) < 0.9
        elif mode == "all_dead":
            dead[:] = True
        else:
            dead = rng.random(m) < 0.1
        keys[dead] = SENT
        pair_key[int(pair_off[v]):int(pair_off[v]) + m] = keys
        live_cnt.append(int((~dead).sum()))
    ent_len = np.array([min(live_cnt[v], visits[v][1]) for v in range(nv)], np.uint32)
    ent_off = np.concatenate([[0], np.cumsum(ent_len)]).astype(np.uint32)
    entries = np.full((int(ent_off[-1]) + 1, 2), 0x77, np.uint64)
    vdone = np.zeros(nv, np.uint8)
    vdone[3] = 1                                                    # a visit the fused kernel already served: untouched
    emu.emu_warp_select(KL, nv, leaf_off.ctypes.data, leaf_len.ctypes.data, members.ctypes.data, ordinals.ctypes.data,
                        vleaf.ctypes.data, vnp.ctypes.data, pair_off.ctypes.data, pair_key.ctypes.data, ent_off.ctypes.data,
                        entries.ctypes.data, vdone.ctypes.data)
    for v, (l, k, mode) in enumerate(visits):
        got = entries[int(ent_off[v]):int(ent_off[v + 1])]
        if vdone[v]:
            assert np.all(got == 0x77)
            continue
        m = leaf_sizes[l]
        keys = pair_key[int(pair_off[v]):int(pair_off[v]) + m]
        ords = ordinals[members[int(leaf_off[l]):int(leaf_off[l]) + m]]
        live = sorted((int(keys[i]), int(ords[i])) for i in range(m) if int(keys[i]) != SENT)
        want = live[:k][: int(ent_len[v])]
        assert [(int(a), int(b)) for a, b in got] == want, (v, l, k, mode)
    assert np.all(entries[-1] == 0x77)                              # nothing written past the last visit
