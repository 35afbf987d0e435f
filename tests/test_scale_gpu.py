"""GPU parity: the resampling kernel through the C ABI vs the CPU oracle (bit-identical to the
reference's STB scaler) and the golden fixtures.  Bit-exact, every geometry."""
import os

import numpy as np
import pytest

import cases
import oracle
from timg_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_scale_cuda_matches_golden_and_oracle(ctx):
    g = np.load(os.path.join(G, "scale.npz"))
    for name, img, ow, oh, fmt in cases.scale_cases():
        got = ctx.scale(img, ow, oh, fmt)
        assert (got == g[name]).all(), f"{name}: vs golden, max diff {np.abs(got.astype(int) - g[name]).max()}"
        assert (got == oracle.stb_resize(img, ow, oh, fmt)).all(), name


def test_scale_cuda_random_geometries(ctx):
    rng = np.random.default_rng(12)
    for it in range(60):
        iw, ih = int(rng.integers(1, 500)), int(rng.integers(1, 400))
        mode = it % 5
        if mode == 0:
            ow, oh = int(rng.integers(1, 500)), int(rng.integers(1, 400))
        elif mode == 1:
            ow, oh = max(1, iw // int(rng.integers(1, 12))), max(1, ih // int(rng.integers(1, 12)))
        elif mode == 2:
            ow, oh = iw * int(rng.integers(1, 4)), ih * int(rng.integers(1, 4))
        elif mode == 3:
            ow, oh = iw, int(rng.integers(1, 400))
        else:
            ow, oh = int(rng.integers(1, 500)), ih
        img = synth.frame_np(900 + it, iw, ih, ["noisea", "photo", "alpha", "noise"][it % 4])
        fmt = it % 2
        want = oracle.stb_resize(img, ow, oh, fmt)
        got = ctx.scale(img, ow, oh, fmt)
        assert (got == want).all(), (iw, ih, ow, oh, fmt, int(np.abs(got.astype(int) - want).max()))


def test_scale_cuda_config_geometries_full_size(ctx):
    """BASELINE configs at full size against the oracle: C3 1080p->320x90 and C2 4K->2700x1519."""
    for iw, ih, ow, oh, kind in [(1920, 1080, 320, 90, "photo"), (3840, 2160, 2700, 1519, "photo"),
                                 (3840, 2160, 337, 190, "noise")]:
        img = synth.frame_np(1234, iw, ih, kind)
        want = oracle.stb_resize(img, ow, oh)
        got = ctx.scale(img, ow, oh)
        assert (got == want).all(), (iw, ih, ow, oh)


def test_identity_scale_is_copy(ctx):
    img = synth.frame_np(5, 333, 77, "noisea")
    assert (ctx.scale(img, 333, 77) == img).all()
