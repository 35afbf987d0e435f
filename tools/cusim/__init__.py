"""cusim: functional simulator of libb200timg's CUDA kernels on the CPU (development / test tool).

`activate()` points the test harness (timg_b200.lib()) at the simulated library so the GPU parity
tests can be replayed without a GPU while a kernel is being written:

    B200TIMG_CUSIM=1 python -m pytest tests/test_sixel_gpu.py -m gpu -x -q

Nothing in the product imports this package; the shipped library has no CPU path.
"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libb200timg_sim.so")


def build():
    subprocess.run([sys.executable, os.path.join(HERE, "build.py")], check=True, stdout=subprocess.DEVNULL)
    return LIB


def activate(rebuild=True):
    import timg_b200
    if rebuild:
        build()
    L = C.CDLL(LIB)
    for name, (res, args) in timg_b200.ABI.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    timg_b200._lib = L
    return L
