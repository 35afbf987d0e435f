"""GPU tests of the geometry passes around the path (timg_b200/csrc/geom.cu): byte moves, exact by construction,
checked against numpy restatements of the reference's loops.
  ApplyExifOp                src/jpeg-source.cc:84-119
  scroll window / crop       src/graphics-magick-source.cc:383-389, :232-237
  trim bounding box          Magick::Image::trim() -- GraphicsMagick is not in the reference tree (PARITY UNPINNED):
                             the documented corner-colour rule with fuzz 0 is what both sides implement.
"""
import numpy as np
import pytest

from timg_b200 import synth

pytestmark = pytest.mark.gpu


def exif_np(fb, mirror, angle):
    """The reference's loops, literally."""
    orig = fb.copy()
    h, w = orig.shape[:2]
    if mirror:
        orig = orig[:, ::-1].copy()                                   # :88-96 each row reversed
    if angle == 180:
        return orig.reshape(-1, 4)[::-1].reshape(h, w, 4).copy()      # :98-104 whole array reversed
    if angle in (90, -90):
        res = np.zeros((w, h, 4), np.uint8)                           # Framebuffer(h, w): width h, height w
        for y in range(h):
            new_x = (h - y - 1) if angle == -90 else y                # :112
            res[:, new_x] = orig[y, :]                                # SetPixel(new_x, x, orig->at(x, y)) for all x
        return res
    return orig


@pytest.mark.parametrize("w,h", [(37, 21), (64, 64), (1, 9), (200, 3)])
def test_exif_ops_equal_the_reference_loops(ctx, w, h):
    fb = synth.frame_np(5 + w, w, h, "noisea")
    for mirror in (False, True):
        for angle in (0, 180, 90, -90):
            assert (ctx.exif_op(fb, mirror, angle) == exif_np(fb, mirror, angle)).all(), (mirror, angle)


def test_scroll_windows_and_crop(ctx):
    w, h, dw, dh = 97, 61, 40, 25
    img = synth.frame_np(3, w, h, "noisea")
    for dx, dy in ((1, 0), (0, 2), (3, 1), (-2, 0), (-1, -1)):
        xs = w if dx and w % abs(dx) else (w // abs(dx) if dx else 1)          # :352-359
        ys = h if dy and h % abs(dy) else (h // abs(dy) if dy else 1)
        cycle = xs * ys // np.gcd(xs, ys)
        x_init = (w - dw - dx * cycle) if dx < 0 else 0                        # :372-375
        y_init = (h - dh - dy * cycle) if dy < 0 else 0
        n = min(cycle + 1, 40)
        got = ctx.windows(img, dw, dh, x_init, y_init, dx, dy, 0, n)
        for k in range(n):
            ys_ = (y_init + dy * k + np.arange(dh)) % h
            xs_ = (x_init + dx * k + np.arange(dw)) % w
            assert (got[k] == img[np.ix_(ys_, xs_)]).all(), (dx, dy, k)
    c = 7                                                                        # --crop-border (:232-237)
    assert (ctx.windows(img, w - 2 * c, h - 2 * c, c, c)[0] == img[c:h - c, c:w - c]).all()


def bbox_np(fb):
    px = fb.view(np.uint32)[..., 0]
    h, w = px.shape
    tl, tr, bl = px[0, 0], px[0, w - 1], px[h - 1, 0]
    ys, xs = np.nonzero(px != tl)
    x1 = np.nonzero((px != tr).any(0))[0]
    y1 = np.nonzero((px != bl).any(1))[0]
    if len(xs) == 0 or len(x1) == 0 or len(y1) == 0 or x1.max() < xs.min() or y1.max() < ys.min():
        return (0, 0, w, h)
    return (int(xs.min()), int(ys.min()), int(x1.max() - xs.min() + 1), int(y1.max() - ys.min() + 1))


def test_trim_bounding_box(ctx):
    rng = np.random.default_rng(4)
    for it in range(12):
        w, h = int(rng.integers(5, 300)), int(rng.integers(5, 200))
        fb = np.zeros((h, w, 4), np.uint8)
        fb[...] = (10, 20, 30, 255)
        if it % 4:
            x0, y0 = int(rng.integers(0, w - 2)), int(rng.integers(0, h - 2))
            x1, y1 = int(rng.integers(x0 + 1, w)), int(rng.integers(y0 + 1, h))
            fb[y0:y1, x0:x1] = synth.frame_np(it, x1 - x0, y1 - y0, "noise")
        assert ctx.trim_bbox(fb) == bbox_np(fb), it
    fb = synth.frame_np(1, 50, 40, "noise")
    assert ctx.trim_bbox(fb) == bbox_np(fb)
