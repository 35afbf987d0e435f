"""GPU parity: AlphaComposeBackground kernel vs oracle and golden (bit-exact)."""
import os

import numpy as np
import pytest

import cases
import oracle
from timg_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_compose_cuda_matches_golden_and_oracle(ctx):
    g = np.load(os.path.join(G, "compose.npz"))
    for name, fb, kw in cases.compose_cases():
        got = ctx.compose_bg(fb, **kw)
        assert (got == g[name]).all(), name
        assert (got == oracle.compose_bg(fb, **kw)).all(), name


def test_compose_cuda_random(ctx):
    rng = np.random.default_rng(19)
    for i in range(20):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        fb = synth.frame_np(700 + i, w, h, "noisea")
        kw = dict(bg=int(rng.integers(0, 2 ** 24)) | 0xff000000, pattern=int(rng.integers(0, 2 ** 32)),
                  pw=int(rng.integers(0, 6)), ph=int(rng.integers(0, 6)), start_row=int(rng.integers(0, h)))
        assert (ctx.compose_bg(fb, **kw) == oracle.compose_bg(fb, **kw)).all(), (i, kw)


def test_all_alpha_times_all_values_exhaustive(ctx):
    """Every (value, alpha) pair against three backgrounds: 3 * 65536 blends, bit-exact."""
    v, a = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    fb = np.stack([v, v[::-1], v.T, a], axis=-1)
    for bg in (oracle.rgba_u32(0, 0, 0), oracle.rgba_u32(255, 255, 255), oracle.rgba_u32(17, 130, 201)):
        assert (ctx.compose_bg(fb, bg) == oracle.compose_bg(fb, bg)).all()


def test_has_transparency(ctx):
    fb = synth.frame_np(1, 50, 40, "photo")
    assert not ctx.has_transparency(fb)
    fb[30, 7, 3] = 254
    assert ctx.has_transparency(fb)
    assert not ctx.has_transparency(fb, start_row=31)
