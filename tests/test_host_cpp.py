"""The C++ host mirror (include/zebra_b200.hpp: Embedding / metric structs / LSHIndex / Database of the reference, stated
in C++ above the C ABI because the image has no Rust toolchain).  tests/cpp/host_mirror_test.cpp is compiled with g++
against libzebra_b200.so and the oracle library; CPU: types, error behaviour (no device -> ZB_ERR_NO_DEVICE, no fallback)
and the store dump crossing between the C++ and Python hosts; GPU: parity of LSHIndex / Metric / Database against the
oracle through the C++ API."""
import os
import subprocess
import uuid

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    from oracle import zb_oracle as zo

    zo.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_test")
    lib, orc = os.path.join(ROOT, "zebra_b200"), os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", out,
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), f"-L{lib}", "-lzebra_b200", f"-L{orc}",
                           "-lzb_oracle", f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{orc}"])
    return out


def test_cpp_host_types_errors_and_store_dump(binary, tmp_path):
    import zebra_b200 as z
    from zebra_b200 import interchange

    r = subprocess.run([binary, "cpu", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all ok" in r.stdout, r.stdout + r.stderr
    # the dump the C++ host wrote, read by the Python host
    dim, zebra, trees, ids, rows = interchange.read_store(str(tmp_path / "cpp.store"))
    assert dim == 4 and rows.tolist() == [[0, 0.5, 1, 1.5], [5, 5.5, 6, 6.5], [10, 10.5, 11, 11.5]]
    assert [int.from_bytes(i.tobytes(), "big") for i in ids] == [1, 2, 3]
    assert interchange.zebra_file_decode(zebra, z.MinkowskiDistance(0)) == (uuid.UUID(bytes=bytes(range(0x10, 0x20))), 3, 5, 2)
    assert [k for k, _ in trees] == [interchange.tree_key(0), interchange.tree_key(1)]
    nodes, coef, cst, leaf_off, mids = interchange.tree_blob_decode(4, trees[0][1])
    assert nodes.tolist() == [[0, 1, 2, -1], [-1, -1, -1, 0], [-1, -1, -1, 1]] and coef.tolist() == [[1, 0, 0, -1]] and cst.tolist() == [0.25]
    assert leaf_off.tolist() == [0, 1, 3] and [int.from_bytes(m.tobytes(), "big") for m in mids] == [1, 2, 3]
    nodes, _, _, leaf_off, mids = interchange.tree_blob_decode(4, trees[1][1])
    assert nodes.tolist() == [[-1, -1, -1, 0]] and leaf_off.tolist() == [0, 3]
    # and the other way round: a dump written by the Python host, read by the C++ host
    p = str(tmp_path / "py.store")
    rng = np.random.default_rng(0)
    prow = rng.standard_normal((5, 6)).astype(np.float32)
    pids = np.arange(80, dtype=np.uint8).reshape(5, 16)
    interchange.write_store(p, 6, b"z" * 40, [(interchange.tree_key(0), b"\x01\0\0\0" + b"\0" * 8)], pids, prow)
    r = subprocess.run([binary, "read", p], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"dim=6 zebra=40 trees=1 rows=5 first_row0={float(prow[0, 0]):g} last_id_byte=79 tree0_bytes=12" in r.stdout


@pytest.mark.gpu
def test_cpp_host_parity_on_gpu(binary, tmp_path):
    r = subprocess.run([binary, "gpu", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "all ok" in r.stdout and "MISMATCH" not in r.stdout and "FAILED" not in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
    assert r.stdout.count(": ok") == 7
