"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/zebra_b200.h declares, and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "zebra_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from zebra_b200 import _ffi

    lib = _ffi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in zebra_b200.h but not exported"
        assert name in _ffi.SYMBOLS, f"{name} has no ctypes prototype"
    exported = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    got = sorted(set(re.findall(r" T (zb_[a-z0-9_]+)", exported)))
    assert got == declared
    assert lib.zb_abi_version() == 2


def test_struct_layouts_match_header():
    from zebra_b200 import _ffi

    assert C.sizeof(_ffi.Options) == 56
    assert _ffi.Options.max_node_size.offset == 8 and _ffi.Options.seed.offset == 24
    assert _ffi.Options.metric_power.offset == 40
    assert C.sizeof(_ffi.Stats) == 13 * 8 + 5 * 4 + 2 * 4 + 8 * 4 + 4  # tail padding to 8


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import zebra_b200 as z

    with pytest.raises(z.ZebraError) as ei:
        z.LSHIndex(384, z.LSHIndexOptions(), z.L2SquaredDistance())
    assert ei.value.code == -3
    with pytest.raises(z.ZebraError):
        z.L2Distance().distance(np.zeros(16, np.float32), np.zeros(16, np.float32))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under zebra_b200/ may reference it."""
    pkg = os.path.join(ROOT, "zebra_b200")
    for dp, _, fns in os.walk(pkg):
        if "build" in dp:
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                assert "zb_oracle" not in src and "from oracle" not in src and "import oracle" not in src, fn


def test_invalid_arguments_are_reported():
    from zebra_b200 import _ffi

    lib = _ffi.lib()
    assert lib.zb_index_create(None, None) == -1
    assert b"NULL" in lib.zb_last_error()
    o = _ffi.Options()
    o.dim, o.metric, o.num_trees = 0, 0, 1
    h = C.c_void_p()
    assert lib.zb_index_create(C.byref(o), C.byref(h)) == -1
    o.dim, o.metric = 16, 12                              # one past the last zb_metric
    assert lib.zb_index_create(C.byref(o), C.byref(h)) == -1
    o.metric, o.metric_power = _ffi.METRIC_MINKOWSKI, 65  # Minkowski power out of range
    assert lib.zb_index_create(C.byref(o), C.byref(h)) == -1 and b"metric_power" in lib.zb_last_error()
    o.metric_power = -1
    assert lib.zb_index_create(C.byref(o), C.byref(h)) == -1
    assert lib.zb_index_destroy(None) == 0
    # every entry point added with ABI v2 checks its arguments before it touches a device
    u = C.c_uint64()
    assert lib.zb_index_load_flat(None, 0, None, None, 4, None, None) == -1
    assert lib.zb_index_export_rows(None, 0, 0, None, None, None) == -1
    assert lib.zb_index_export_tree_blob(None, 0, None, 0, C.byref(u)) == -1
    assert lib.zb_index_export_tree_blobs(None, None, 0, None, C.byref(u)) == -1
    assert lib.zb_index_options(None, None) == -1
    assert lib.zb_tree_blob_decode(0, None, 0, None, None, None, None, None, None) == -1
    assert lib.zb_store_flatten(4, 0, None, 0, None, None, None, None) == -1
    assert lib.zb_flat_store_free(None) == 0
    assert lib.zb_zebra_file_encode(None, 0, 0, 5, 15, None, 0, C.byref(u)) == -1
    assert lib.zb_metric_distance_batch(0, 12, 0, 0, 16, None, None, None) == -1          # not a zb_metric
    assert lib.zb_metric_distance_batch(0, 10, 65, 0, 16, None, None, None) == -1         # Minkowski power out of range


def test_stats_struct_offsets_match_the_c_compiler(tmp_path):
    """Every field of zb_stats (round 2 gave the reserved words to the dot-product filter's counters) sits where the C compiler
    puts it: sizeof and offsetof of the header against the ctypes mirror the tests and the bench read."""
    from zebra_b200 import _ffi

    names = [n for n, _ in _ffi.Stats._fields_]
    body = "".join(f'  printf("%s %zu\\n", "{n}", offsetof(zb_stats, {n}));\n' for n in names)
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "zebra_b200.h"\n'
                   'int main(void) {\n  printf("sizeof %zu\\n", sizeof(zb_stats));\n' + body + '  return 0; }\n')
    exe = str(tmp_path / "off")
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), "-o", exe, str(src)])
    out = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    assert int(out.pop("sizeof")) == C.sizeof(_ffi.Stats)
    for n in names:
        assert int(out[n]) == getattr(_ffi.Stats, n).offset, n


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/zebra_b200.h must compile as C99 (no C++-isms), and a C program links against it."""
    src = tmp_path / "abi.c"
    src.write_text('#include "zebra_b200.h"\n'
                   'int main(void) { zb_options o; zb_stats s; zb_import_report r; int n = -1; (void)o; (void)s; (void)r;\n'
                   '  if (zb_abi_version() != ZB_ABI_VERSION) return 1;\n'
                   '  if (zb_device_count(&n) != ZB_OK || n < 0) return 2;\n'
                   '  if (zb_index_create(0, 0) != ZB_ERR_INVALID || !zb_last_error()[0]) return 3;\n'
                   '  return 0; }\n')
    exe = str(tmp_path / "abi_c")
    lib = os.path.join(ROOT, "zebra_b200")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           str(src), "-L" + lib, "-lzebra_b200", "-Wl,-rpath," + lib])
    assert subprocess.run([exe]).returncode == 0
