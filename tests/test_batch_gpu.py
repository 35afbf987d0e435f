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


@pytest.mark.parametrize("chunk", [None, "2", "1", "4"])
def test_blocks_batch_animation_delta_matches_canvas_sequence(ctx, chunk, monkeypatch):
    """animation=1: frame f is delta-encoded against frame f-1, as a canvas receiving Send(dy=-h); the host pipeline
    may cut the animation into chunks (each later chunk re-uploads one halo frame)."""
    if chunk:
        monkeypatch.setenv("B200TIMG_CHUNK_FRAMES", chunk)
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


def test_blocks_animation_sharded_with_halo_frames_equals_unsharded(ctx):
    """a rank that owns frames [lo, hi) of a delta-encoded animation loads the halo frame lo-1 and encodes
    with animation = 2; the concatenation over ranks is byte-identical to the unsharded run."""
    from timg_b200 import shard
    n, w, h = 11, 128, 64
    frames = []
    for k in range(n):
        fr = synth.frame_np(70, w, h, "photo")
        fr[4 + 3 * k:12 + 3 * k, 10 + 9 * k:18 + 9 * k] = synth.frame_np(80 + k, 8, 8, "noise")
        frames.append(fr)
    frames = np.stack(frames)
    whole = ctx.blocks_batch(frames, _batch(n, w, h, w, h, flags=timg_b200.QUARTER, animation=1))
    for world in (2, 3, 4):
        got = []
        for r in range(world):
            first, cnt, anim = shard.animation_chunk(n, r, world)
            outs = ctx.blocks_batch(frames[first:first + cnt], _batch(cnt, w, h, w, h, flags=timg_b200.QUARTER, animation=anim))
            if anim == 2:
                assert outs[0] == b""
                outs = outs[1:]
            got += outs
        assert got == whole, world


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


def test_sixel_batch_c5_shape_unscaled_720p(ctx):
    """C5 geometry: 1280x720 frames shown unscaled (copy-only scaler, height already a multiple of 6)."""
    n, w, h = 3, 1280, 720
    assert timg_b200.calc_fit(w, h, 2700, 1800, 9, 18) == (False, w, h)
    frames = np.stack([synth.frame_np(200 + i, w, h, "photo") for i in range(n)])
    outs = ctx.sixel_batch(frames, _batch(n, w, h, w, h))
    for f in range(n):
        img, used = oracle.sixel_decode(outs[f])
        want, _ = oracle.sixel_decode(oracle.sixel_encode(frames[f], mode=1))
        assert img.shape == (h, w, 3) and used <= 256
        assert (img == want).all(), f


def test_sixel_device_entry_matches_host_entry(ctx):
    """b200timg_sixel_dev on device-resident frames == b200timg_sixel_encode frame by frame."""
    import ctypes as C
    import torch
    n, w, h = 3, 200, 96
    frames = np.stack([synth.frame_np(300 + i, w, h, "photo") for i in range(n)])
    d = torch.tensor(frames).cuda()
    out = torch.zeros(n * (4096 + 6 * w * h), dtype=torch.uint8, device="cuda")
    offs = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    rc = timg_b200.lib().b200timg_sixel_dev(ctx.h, d.data_ptr(), w, h, n, out.data_ptr(), out.numel(), offs.data_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    o, ob = offs.cpu().numpy(), out.cpu().numpy()
    for f in range(n):
        assert ob[o[f]:o[f + 1]].tobytes() == ctx.sixel_encode(frames[f]), f


def test_blocks_full_size_properties_4k_one_to_one(ctx):
    """4K quarter-block frame (1:1): size-independent properties instead of an oracle run:
    every row pair ends with ESC[0m LF, and re-encoding the same frame as a delta is empty."""
    fb = synth.frame_np(9, 3840, 2160, "photo")
    cv = B200BlockCanvas(ctx, True)
    full = cv.send(fb)
    assert full.count(b"\033[0m\n") == 1080
    assert cv.send(fb, 0, -2160) == b""
    fb2 = fb.copy()
    fb2[1001, 2001] = (1, 2, 3, 255)
    delta = cv.send(fb2, 0, -2160)
    assert 0 < len(delta) < 200 and b"\033[1000C" in delta
