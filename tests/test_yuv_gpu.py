"""GPU tests of the libswscale-side scalers (bilinear.cu): the video front end (decoder YUV 4:2:0 -> RGBA at the
target size, src/video-source.cc:59-89,352-354) and the default build's SWS_BILINEAR ImageScaler
(src/image-scaler.cc:45-72).

PARITY UNPINNED: libswscale is not part of the reference tree and no version is pinned.  Two checks:
  * against the float64 numpy statement of the same filter (oracle.yuv420_to_rgba_np / bilinear_rgba_np): <= 1 LSB;
  * against the libswscale 9.1 that happens to ship inside the image's OpenCV wheel, called exactly the way the
    reference calls it (sws_getContext(..., SWS_BILINEAR) + sws_scale): stated distance bounds below.  Measured:
    RGBA -> RGBA max 1 / mean 0.02-0.05; YUV420P -> RGBA max 3-10 / mean 0.7-2.5 (libswscale's fixed-point
    yuv2rgb tables and 15-bit intermediates account for the rest).
"""
import numpy as np
import pytest

import oracle
import timg_b200
from timg_b200 import synth

pytestmark = pytest.mark.gpu

SWS_RGBA_MAX, SWS_RGBA_MEAN = 2, 0.25          # stated tolerance vs libswscale, RGBA -> RGBA
SWS_YUV_MAX, SWS_YUV_MEAN = 12, 3.0            # stated tolerance vs libswscale, YUV420P -> RGBA

GEOMS = [(640, 480, 450, 337, "photo"), (1920, 1080, 320, 90, "photo"), (256, 128, 300, 200, "noise"),
         (3840, 216, 2700, 152, "photo"), (64, 48, 64, 48, "noise"), (130, 98, 67, 50, "alpha")]


def _planes(yuv, iw, ih):
    cw, ch = iw // 2, ih // 2
    return (yuv[: iw * ih].reshape(ih, iw), yuv[iw * ih: iw * ih + cw * ch].reshape(ch, cw),
            yuv[iw * ih + cw * ch:].reshape(ch, cw))


@pytest.mark.parametrize("iw,ih,ow,oh,kind", GEOMS)
def test_yuv420_to_rgba_matches_restatement_and_libswscale(ctx, iw, ih, ow, oh, kind):
    yuv = oracle.rgba_to_i420_np(synth.frame_np(3 + iw, iw, ih, kind))
    got = ctx.yuv_scale(yuv, iw, ih, ow, oh)
    want = oracle.yuv420_to_rgba_np(yuv, iw, ih, ow, oh)
    assert np.abs(got.astype(int) - want).max() <= 1
    assert (got[..., 3] == 255).all()
    # NV12 carries the same samples interleaved
    Y, U, V = _planes(yuv, iw, ih)
    nv12 = np.concatenate([Y.reshape(-1), np.stack([U, V], -1).reshape(-1)])
    assert (ctx.yuv_scale(nv12, iw, ih, ow, oh, timg_b200.FMT_NV12) == got).all()
    # full range (the reference's YUVJ formats)
    gotj = ctx.yuv_scale(yuv, iw, ih, ow, oh, timg_b200.FMT_I420 | timg_b200.FMT_FULL_RANGE)
    assert np.abs(gotj.astype(int) - oracle.yuv420_to_rgba_np(yuv, iw, ih, ow, oh, full_range=True)).max() <= 1
    if oracle.swscale():
        ref = oracle.sws_scale_np([Y, U, V], [iw, iw // 2, iw // 2], iw, ih, 0, ow, oh)
        e = np.abs(got[..., :3].astype(int) - ref[..., :3])
        assert e.max() <= SWS_YUV_MAX and e.mean() <= SWS_YUV_MEAN, (int(e.max()), float(e.mean()))


@pytest.mark.parametrize("iw,ih,ow,oh,kind", GEOMS)
def test_bilinear_rgba_matches_restatement_and_libswscale(ctx, iw, ih, ow, oh, kind):
    img = synth.frame_np(9 + iw, iw, ih, kind)
    img[..., 3] = 255
    for fmt in (0, 1):
        got = ctx.scale(img, ow, oh, fmt, fast=2)
        assert np.abs(got.astype(int) - oracle.bilinear_rgba_np(img, ow, oh, fmt)).max() <= 1
    if oracle.swscale():
        got = ctx.scale(img, ow, oh, 0, fast=2)
        ref = oracle.sws_scale_np([img], [iw * 4], iw, ih, 26, ow, oh)
        e = np.abs(got[..., :3].astype(int) - ref[..., :3])
        assert e.max() <= SWS_RGBA_MAX and e.mean() <= SWS_RGBA_MEAN, (int(e.max()), float(e.mean()))


def test_yuv_batches_equal_staged_pipeline(ctx):
    """I420 frames through the batch entry points == b200timg_yuv_scale followed by the RGBA stages."""
    n, iw, ih = 3, 320, 240
    _, ow, oh = timg_b200.calc_fit(iw, ih, 80, 50, 1, 2)
    yuvs = np.stack([oracle.rgba_to_i420_np(synth.frame_np(70 + i, iw, ih, "photo")) for i in range(n)])
    b = timg_b200.Batch(n_frames=n, src_w=iw, src_h=ih, src_fmt=timg_b200.FMT_I420, out_w=ow, out_h=oh, has_bg=1,
                        bg=timg_b200.rgba_u32(0, 0, 0), pattern=0, pattern_w=0, pattern_h=0, flags=0, x_indent_cells=0, animation=0)
    blocks = ctx.blocks_batch(yuvs, b)
    for f in range(n):
        fb = ctx.yuv_scale(yuvs[f], iw, ih, ow, oh)
        assert blocks[f] == ctx.blocks_encode(fb), f
    ow2, oh2 = 200, 150
    b2 = timg_b200.Batch.from_buffer_copy(b)
    b2.out_w, b2.out_h = ow2, oh2
    sixels = ctx.sixel_batch(yuvs, b2)
    for f in range(n):
        fb = ctx.yuv_scale(yuvs[f], iw, ih, ow2, oh2)
        assert sixels[f] == ctx.sixel_encode(fb), f
