"""GPU parity for the sixel path.  libsixel is not vendored in the reference (PARITY UNPINNED), so:
  * against the CPU restatement with the same order-free semantics (oracle mode 1) the device must be
    BIT-IDENTICAL: palette, index plane, decoded image;
  * against the libsixel-faithful restatement (oracle mode 0: raster order, first-come memo) it must
    be within a stated tolerance: mean CIE76 delta-E after a 5x5 box blur < 3.0 between the two
    decoded images, and its error against the source within 0.5 delta-E of the restatement's.
"""
import re

import numpy as np
import pytest

import oracle
import timg_b200
from timg_b200 import synth

pytestmark = pytest.mark.gpu

DE_BETWEEN = 3.0     # stated tolerance: device vs libsixel-faithful restatement (mean delta-E after a 5x5 box blur)
DE_SOURCE = 0.5      # stated tolerance: |error vs source (device) - error vs source (restatement)|, blurred
# The same WITHOUT the blur.  The two order-free substitutions (sixel.cu header) change the palette (tie order in
# the median cut) and with it nearly every pixel's index: measured 90-98 % of the pixels decode to a different
# colour than the libsixel-faithful restatement's, mean delta-E between the two images 6-12 (two different dither
# patterns of the same picture).  What is preserved is the quality: the error against the SOURCE is the same.
DE_SOURCE_UNBLURRED = 1.5    # |mean dE(device, source) - mean dE(restatement, source)|, no blur (measured 0.05-1.05)
DE_BETWEEN_UNBLURRED = 14.0  # mean dE(device, restatement), no blur (measured 6.4-12.2; informational upper bound)


def pct(pal):
    return ((pal.astype(int) * 100 + 127) // 255) * 255 // 100


CASES = [("photo", 337, 192), ("noise", 337, 192), ("alpha", 160, 120), ("photo", 64, 6), ("photo", 1, 6),
         ("photo", 2, 12), ("noise", 31, 36), ("photo", 33, 66), ("noise", 65, 6), ("photo", 675, 384),
         ("noisea", 129, 96)]


@pytest.mark.parametrize("kind,w,h", CASES)
def test_sixel_cuda_bit_identical_to_device_semantics_oracle(ctx, kind, w, h):
    fb = synth.frame_np(4321 + w, w, h, kind)
    data = ctx.sixel_encode(fb)
    pal, orig, idx = ctx.sixel_debug(w, h)
    _, det = oracle.sixel_encode(fb, True, mode=1)
    assert orig == det["origcolors"]
    assert pal.shape == det["palette"].shape and (pal == det["palette"]).all(), "palette"
    assert (idx == det["index"]).all(), f"index plane differs in {(idx != det['index']).sum()} px"
    img, used = oracle.sixel_decode(data)
    assert img.shape == (h, w, 3)
    assert (img == pct(pal)[idx]).all(), "stream does not decode to palette[index]"
    assert data.startswith(b'\x1bPq"1;1;%d;%d#0;2;' % (w, h)) and data.endswith(b"\x1b\\")
    assert re.fullmatch(rb'\x1bPq"1;1;\d+;\d+(#\d+;2;\d+;\d+;\d+)+[#!$\-0-9?-~]+\x1b\\', data)


@pytest.mark.parametrize("kind,w,h", [("photo", 337, 192), ("noise", 200, 96), ("alpha", 160, 120)])
def test_sixel_cuda_within_tolerance_of_libsixel_semantics(ctx, kind, w, h):
    fb = synth.frame_np(1234, w, h, kind)
    got, _ = oracle.sixel_decode(ctx.sixel_encode(fb))
    ref, _ = oracle.sixel_decode(oracle.sixel_encode(fb, mode=0))
    src = fb[..., :3]
    assert oracle.mean_delta_e(got, ref, 2) < DE_BETWEEN
    assert abs(oracle.mean_delta_e(got, src, 2) - oracle.mean_delta_e(ref, src, 2)) < DE_SOURCE
    assert abs(oracle.mean_delta_e(got, src, 0) - oracle.mean_delta_e(ref, src, 0)) < DE_SOURCE_UNBLURRED
    assert oracle.mean_delta_e(got, ref, 0) < DE_BETWEEN_UNBLURRED


def test_sixel_cuda_against_real_libsixel_if_the_host_has_it(ctx):
    """SURVEY 8c acceptance item 4: libsixel is not in the reference tree or the image, but if this host has one
    (find_library / pkg-config), the device stream is compared with the real library run through the reference's
    exact call sequence (oracle/libsixel_probe.py).  Skipped, with the probe's log as the reason, otherwise."""
    from oracle import libsixel_probe
    lib, how = libsixel_probe.find()
    if lib is None:
        pytest.skip("no libsixel on this host: " + how)
    for kind, w, h in [("photo", 337, 192), ("noise", 200, 96)]:
        fb = synth.frame_np(1234, w, h, kind)
        got, _ = oracle.sixel_decode(ctx.sixel_encode(fb))
        ref, _ = oracle.sixel_decode(libsixel_probe.encode(fb))
        src = fb[..., :3]
        assert oracle.mean_delta_e(got, ref, 2) < DE_BETWEEN
        assert abs(oracle.mean_delta_e(got, src, 0) - oracle.mean_delta_e(ref, src, 0)) < DE_SOURCE_UNBLURRED


def test_sixel_cuda_few_colours_no_dither(ctx):
    fb = np.zeros((12, 40, 4), np.uint8)
    fb[..., 3] = 255
    fb[:, :20, 0] = 200
    fb[6:, :, 1] = 96
    img, used = oracle.sixel_decode(ctx.sixel_encode(fb))
    assert used == 4 and (img == pct(fb[..., :3] & 0xF8)).all()


def test_sixel_cuda_long_runs_and_solid(ctx):
    fb = np.zeros((18, 500, 4), np.uint8)
    fb[..., :3] = (40, 80, 120)
    fb[..., 3] = 255
    data = ctx.sixel_encode(fb)
    img, used = oracle.sixel_decode(data)
    assert used == 1 and (img == pct(np.array([40, 80, 120]) & 0xF8)).all()
    assert b"!500~" in data                                  # one RLE run per band


def test_sixel_cuda_c4_shape_full_pipeline_matches_staged(ctx):
    """C4 geometry: 4K -> 337x190 -> pad 192 -> sixel, as one batch call vs stage by stage."""
    n, iw, ih = 2, 3840, 2160
    frames = synth.frames_np(99, n, iw, ih, "photo")
    _, ow, oh = timg_b200.calc_fit(iw, ih, 337, 225, 9, 18)
    assert (ow, oh) == (337, 190)
    b = timg_b200.Batch(n_frames=n, src_w=iw, src_h=ih, src_fmt=0, out_w=ow, out_h=oh, has_bg=1,
                        bg=timg_b200.rgba_u32(10, 20, 30), pattern=0, pattern_w=0, pattern_h=0, flags=0,
                        x_indent_cells=0, animation=0)
    outs = ctx.sixel_batch(frames, b)
    for f in range(n):
        fbs = ctx.scale(frames[f], ow, oh)
        padded = np.zeros((192, ow, 4), np.uint8)
        padded[:oh] = fbs
        padded = ctx.compose_bg(padded, timg_b200.rgba_u32(10, 20, 30))
        assert outs[f] == ctx.sixel_encode(padded)
        assert (padded[oh:, :, :3] == (10, 20, 30)).all()


def test_sixel_cuda_c2_full_size_decodes(ctx):
    """C2 at full size (2700x1524): bit-identical index plane vs the oracle, stream decodes."""
    fb = synth.frame_np(1234, 2700, 1524, "photo")
    data = ctx.sixel_encode(fb)
    pal, orig, idx = ctx.sixel_debug(2700, 1524)
    _, det = oracle.sixel_encode(fb, True, mode=1)
    assert (pal == det["palette"]).all()
    assert (idx == det["index"]).all()
    img, _ = oracle.sixel_decode(data)
    assert (img == pct(pal)[idx]).all()
