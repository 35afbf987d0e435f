"""The C++ adapters (timg_b200/csrc/adapters.h) must compile against the reference's own headers.
Only possible where the reference tree is mounted (the build container)."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "terminal-canvas.h")), reason="reference tree absent")
def test_adapters_compile_against_reference_headers():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "check.cc")
        open(src, "w").write('#include "adapters.h"\nint main() { return 0; }\n')
        open(os.path.join(d, "timg-version.h"), "w").write('#define TIMG_VERSION "check"\n')
        r = subprocess.run(["g++", "-std=gnu++17", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter",
                            "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "timg_b200", "csrc"),
                            "-I", REF, "-I", d, src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
