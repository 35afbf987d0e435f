"""A handful of small calls through every kernel family, meant to be run under compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import timg_b200  # noqa: E402
from timg_b200 import synth  # noqa: E402

ctx = timg_b200.Context(0)
for (iw, ih, ow, oh, kind) in [(128, 96, 90, 67, "photo"), (128, 96, 90, 67, "noisea"), (200, 150, 40, 30, "alpha"),
                               (64, 48, 160, 100, "noise"), (77, 33, 77, 33, "noisea"), (131, 70, 100, 60, "alpha")]:
    out = ctx.scale(synth.frame_np(1, iw, ih, kind), ow, oh)
    assert out.shape == (oh, ow, 4)
fb = synth.frame_np(2, 66, 41, "alpha")
ctx.compose_bg(fb, timg_b200.rgba_u32(10, 20, 30), timg_b200.rgba_u32(90, 90, 90), 4, 2, 3)
ctx.has_transparency(fb)
for flags in (0, timg_b200.QUARTER, timg_b200.UPPER | timg_b200.COLOR8):
    w = 64
    a, b = synth.frame_np(3, w, 41, "noisea"), synth.frame_np(4, w, 41, "noisea")
    ctx.blocks_encode(a, None, flags, 2)
    ctx.blocks_encode(b, a, flags, 2)
for (w, h, kind) in [(96, 48, "photo"), (33, 6, "noise"), (130, 72, "alpha")]:
    fb = synth.frame_np(5, w, h, kind)
    fb[..., 3] = 255
    assert len(ctx.sixel_encode(fb)) > 10
# batched host pipelines (3 streams, 2 chunks each), incl. delta-coded animation and a window with holes
def batch(n, iw, ih, ow, oh, **kw):
    d = dict(n_frames=n, src_w=iw, src_h=ih, src_fmt=0, out_w=ow, out_h=oh, has_bg=1, bg=timg_b200.rgba_u32(10, 20, 30),
             pattern=0, pattern_w=0, pattern_h=0, flags=0, x_indent_cells=0, animation=0)
    d.update(kw)
    return timg_b200.Batch(**d)


os.environ["B200TIMG_CHUNK_FRAMES"] = "2"
frames = np.stack([synth.frame_np(10 + i, 128, 96, "alpha" if i & 1 else "photo") for i in range(3)])
frames[0, 10:40, 20:70, 3] = 0
assert len(ctx.sixel_batch(frames, batch(3, 128, 96, 90, 67))) == 3
assert len(ctx.blocks_batch(frames, batch(3, 128, 96, 60, 44, flags=timg_b200.QUARTER, animation=1))) == 3
print("sanitize cases done")
