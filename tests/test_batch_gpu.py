"""Batched host entry points (scale -> compose -> encode for many frames per call) against the
single-frame path and the oracle; also exercises the chunked upload/compute/download pipeline."""
import os

import numpy as np
import pytest

import oracle
import timg_b200
from timg_b200 import synth
from timg_b200.canvas import B200BlockCanvas

pytestmark = pytest.mark.gpu


def _batch(n, iw, ih, ow, oh, **kw):
    d = dict(n_frames=n, src_w=iw, src_h=ih, src_fmt=0, out_w=ow, out_h=oh, has_bg=1, bg=timg_b200.rgba_u32(0, 0, 0),
             pattern=0, pattern_w=0, pattern_h=0, flags=0, x_indent_cells=0, animation=0)
    d.update(kw)
    return timg_b200.Batch(**d)


@pytest.mark.parametrize("chunk", [None, "2", "1"])
def test_blocks_batch_c3_shape_matches_oracle(ctx, chunk, monkeypatch):
    """C3 geometry: 1080p frames -> 320x90 -> -p quarter, distinct frames, full emission each."""
    if chunk:
        monkeypatch.setenv("B200TIMG_CHUNK_FRAMES", chunk)
    n, iw, ih = 5, 1920, 1080
    _, ow, oh = timg_b200.calc_fit(iw, ih, 320, 100, 2, 2, 2.0)
    assert (ow, oh) == (320, 90)
    frames = np.stack([synth.frame_np(50 + i, iw, ih, "alpha" if i % 2 else "photo") for i in range(n)])
    outs = ctx.blocks_batch(frames, _batch(n, iw, ih, ow, oh, flags=timg_b200.QUARTER, x_indent_cells=3))
    for f in range(n):
        fb = oracle.compose_bg(oracle.stb_resize(frames[f], ow, oh), oracle.rgba_u32(0, 0, 0))
        assert outs[f] == oracle.BlockCanvas(True).send(fb, x=6), f


def test_blocks_batch_animation_delta_matches_canvas_sequence(ctx):
    """animation=1: frame f is delta-encoded against frame f-1, as a canvas receiving Send(dy=-h)."""
    n, w, h = 6, 128, 64
    frames = []
    for k in range(n):
        fr = synth.frame_np(70, w, h, "photo")
        fr[4 + 3 * k:12 + 3 * k, 10 + 9 * k:18 + 9 * k] = synth.frame_np(80 + k, 8, 8, "noise")
        frames.append(fr)
    frames = np.stack(frames)
    outs = ctx.blocks_batch(frames, _batch(n, w, h, w, h, animation=1))
    cv = oracle.BlockCanvas(False)
    for f in range(n):
        want = cv.send(frames[f], 0, 0 if f == 0 else -h)
        prefix = b"" if f == 0 else b"\033[%dA" % (h // 2)      # the adapter adds the cursor-up, the ABI returns image bytes
        assert prefix + outs[f] == want, f


@pytest.mark.parametrize("chunk", ["2", None])
def test_sixel_batch_chunked_equals_single_frames(ctx, chunk, monkeypatch):
    if chunk:
        monkeypatch.setenv("B200TIMG_CHUNK_FRAMES", chunk)
    n, iw, ih, ow, oh = 5, 400, 300, 200, 150
    frames = np.stack([synth.frame_np(90 + i, iw, ih, "photo") for i in range(n)])
    outs = ctx.sixel_batch(frames, _batch(n, iw, ih, ow, oh))
    for f in range(n):
        fbs = ctx.scale(frames[f], ow, oh)
        assert outs[f] == ctx.sixel_encode(fbs), f


def test_sixel_batch_too_small_buffer_reports_enospc(ctx):
    import ctypes as C
    n, w, h = 2, 120, 60
    frames = np.stack([synth.frame_np(3 + i, w, h, "photo") for i in range(n)])
    out = np.empty(100, np.uint8)
    offs = np.zeros(n + 1, np.uint64)
    b = _batch(n, w, h, w, h)
    rc = timg_b200.lib().b200timg_sixel_batch(ctx.h, C.byref(b), frames.ctypes.data, out.ctypes.data, out.size,
                                               offs.ctypes.data)
    assert rc == timg_b200.ENOSPC
