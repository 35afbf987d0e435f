"""Build tools/cusim/_build/libb200timg_sim.so: the library's .cu sources compiled by g++ against the
cusim shim (see cuda_runtime.h here).  DEVELOPMENT / TEST TOOL ONLY -- a functional simulator to debug
kernel logic against the oracle on a box without a GPU.  Never part of the product library.

    python tools/cusim/build.py            # build (incremental)
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "timg_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libb200timg_sim.so")

LAUNCH = re.compile(r"([A-Za-z_][\w:.]*(?:->\w+)*(?:<[^<>();]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
EXT_SHARED = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:]+)\s+(\w+)\s*\[\s*\]\s*;")


def rewrite(text):
    out, pos = [], 0
    for m in LAUNCH.finditer(text):
        out.append(text[pos:m.start()])
        rest = text[m.end():].lstrip()
        sep = "" if rest.startswith(")") else ", "
        out.append(f"cusim::Launcher{{{m.group(2)}}}.run({m.group(1)}{sep}")
        pos = m.end()
    out.append(text[pos:])
    return re.sub(r"extern\s+__shared__", "extern", "".join(out))


def main():
    os.makedirs(OUT, exist_ok=True)
    cus = sorted(f for f in os.listdir(SRC) if f.endswith(".cu"))
    shared = {}
    gen = []
    for f in cus:
        text = open(os.path.join(SRC, f)).read()
        for ty, name in EXT_SHARED.findall(text):
            if shared.setdefault(name, ty) != ty:
                sys.exit(f"cusim: dynamic shared array {name} declared with two types")
        dst = os.path.join(OUT, f[:-3] + ".sim.cc")
        new = f'#line 1 "{os.path.join(SRC, f)}"\n' + rewrite(text)
        if not os.path.exists(dst) or open(dst).read() != new:
            open(dst, "w").write(new)
        gen.append(dst)
    sh = os.path.join(OUT, "shared_arrays.sim.cc")
    body = '#include "cuda_runtime.h"\n#include "common.cuh"\nnamespace b200timg {\n' + "".join(
        f"alignas(128) {ty} {name}[(232448 + sizeof({ty}) - 1) / sizeof({ty})];\n" for name, ty in sorted(shared.items())) + "}\n"
    if not os.path.exists(sh) or open(sh).read() != body:
        open(sh, "w").write(body)
    gen.append(sh)
    gen.append(os.path.join(HERE, "cusim.cc"))
    flags = ["-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-x", "c++",
             "-I", HERE, "-I", SRC, "-I", os.path.join(ROOT, "include")]
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "cuda_runtime.h"), os.path.join(ROOT, "include", "b200timg.h")]
    newest_dep = max(os.path.getmtime(d) for d in deps)

    def compile_one(src):
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_dep):
            return obj
        r = subprocess.run(["g++", *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stderr[-6000:])
            sys.exit(f"cusim: compiling {src} failed")
        return obj

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(compile_one, gen))
    if not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        subprocess.run(["g++", "-shared", "-o", LIB, *objs, "-lpthread"], check=True)
    return LIB


if __name__ == "__main__":
    print(main())
