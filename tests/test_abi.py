"""The C-ABI library loads, exports every symbol include/b200timg.h declares, and fails
loudly without a GPU (no compute calls here)."""
import os
import re

import pytest

import timg_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200timg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200timg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    L = timg_b200.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200timg.h but not exported"
        assert n in timg_b200.ABI, f"{n} has no ctypes signature in timg_b200.ABI"
    assert set(timg_b200.ABI) == set(names)


def test_bounds_are_host_only():
    L = timg_b200.lib()
    assert L.b200timg_blocks_bound(67, 50) >= 25 * (67 * 39 + 5)
    assert L.b200timg_sixel_bound(2700, 1524) > 2700 * 1524


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(timg_b200.B200Error) as e:
        timg_b200.Context(0)
    assert e.value.code == timg_b200.ENODEV


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ anywhere."""
    pkg = os.path.join(ROOT, "timg_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "libtimg_ref" not in txt, fn


def test_header_is_plain_c(tmp_path):
    """include/b200timg.h is the FFI contract: it must compile as C99 (no C++/torch types), and a C program
    must link against the library with nothing but the header."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200timg.h"\n'
                   'int main(void) { b200timg_fit_opts o; (void)o; return b200timg_version() > 0 && '
                   'b200timg_blocks_bound(80, 50) > 0 && b200timg_as256(0xff102030u) >= 16 ? 0 : 1; }\n')
    exe = tmp_path / "hdr"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        str(src), "-L", os.path.join(root, "timg_b200"), "-lb200timg",
                        "-Wl,-rpath," + os.path.join(root, "timg_b200"), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0      # host-only entry points: no GPU needed
