"""CPU: the C-ABI shared library loads, exports every symbol include/pulpo_b200.h declares, the
ctypes table mirrors the header, and argument validation works without a GPU (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pulpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pulpo_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pulpo_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    from pulpo_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(raw, n), "libpulpo_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_header_arity_matches_ctypes_table():
    from pulpo_b200 import _lib
    src = open(os.path.join(ROOT, "include", "pulpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), "%s: header has %d params, ctypes table %d" % (name, n, len(args))


def test_version_and_errors(lib):
    assert lib.pulpo_version() == 100
    assert lib.pulpo_strerror(0) == b"ok"
    # argument validation happens before any CUDA call -> safe without a GPU
    assert lib.pulpo_warp3d_fwd(None, None, None, None, 1, 1, 4, 4, 4, 0, None) == -1          # null pointer
    one = ctypes.c_void_p(16)
    assert lib.pulpo_warp3d_fwd(one, one, one, None, 1, 1, 1, 4, 4, 0, None) == -2             # S == 1 axis
    assert lib.pulpo_warp3d_fwd(one, one, one, None, 1, 1, 4, 4, 4, 7, None) == -3             # bad coord mode
    assert lib.pulpo_ncc_fwd(one, one, one, None, one, 1 << 20, 4, 0.05, 1, 1, 8, 8, 8, None) == -3   # even window
    assert lib.pulpo_ncc_fwd(one, one, one, None, one, 8, 9, 0.05, 1, 1, 8, 8, 8, None) == -4   # workspace too small
    assert lib.pulpo_resize_up_fwd(one, None, one, 1, 1.0, 1, 3, 4, 4, 4, None) == -3           # factor < 2
    assert lib.pulpo_vecint_fwd(one, one, one, 0, 7, 1, 1, 4, 4, 4, 0, None) == -4              # workspace too small
    assert b"workspace" in lib.pulpo_strerror(-4)


def test_workspace_arithmetic(lib):
    # 7 saved float4 states + the synchronisation tail (row counters, per-CTA maxima): < 64 KB on top
    assert 0 < lib.pulpo_vecint_ws_bytes(7, 1, 1, 80, 96, 112) - 7 * 80 * 96 * 112 * 16 <= 65536
    assert lib.pulpo_vecint_ws_bytes(7, 0, 2, 8, 8, 8) == 2 * 2 * 512 * 16          # forward only: two ping-pong states
    assert 0 < lib.pulpo_vecint_bwd_scratch_bytes(1, 8, 8, 8) - 3 * 512 * 16 <= 65536
    assert lib.pulpo_ncc_ws_bytes(1, 1, 160, 192, 224) >= 16 + 8 * 7 * 12
    assert lib.pulpo_reduce_ws_bytes() >= 16 + 8 * 592


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (that would void every parity claim)."""
    pkg = os.path.join(ROOT, "pulpo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/pulpo_b200.h compiles as pedantic C99 (no C++, no torch types) and a
    plain C program links against the shared library and calls it (no compute call without a GPU)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from pulpo_b200 import build as b
    src = tmp_path / "use_abi.c"
    src.write_text('#include "pulpo_b200.h"\n#include <string.h>\n'
                   'int main(void) { return (pulpo_version() >= 100 && strcmp(pulpo_strerror(0), "ok") == 0 &&\n'
                   '                         pulpo_warp3d_fwd(0, 0, 0, 0, 1, 1, 4, 4, 4, 0, 0) == PULPO_ERR_NULL_POINTER) ? 0 : 1; }\n')
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(b.LIB)
    exe = tmp_path / "use_abi"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-lpulpo_b200", "-Wl,-rpath," + libdir])
    assert subprocess.run([str(exe)]).returncode == 0
