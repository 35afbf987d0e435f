"""Randomised parity sweep on the GPU (beyond the fixed cases in tests/): resampler vs the STB restatement
on random geometries biased towards the planar kernel's class, and sixel vs the mode-1 restatement on
random small frames.  python tools/parity_sweep.py [seed] [n_scale] [n_sixel]   (exit code 1 on a mismatch)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle  # noqa: E402
import timg_b200  # noqa: E402
from timg_b200 import synth  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_scale = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n_sixel = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rng = np.random.default_rng(seed)
if os.environ.get("B200TIMG_CUSIM"):          # developer aid: run the sweep on the CPU simulator of the kernels
    from tools import cusim
    cusim.activate()
ctx = timg_b200.Context(0)
bad = 0
kernels = {}
for it in range(n_scale):
    iw = int(rng.integers(1, 260)) * 4 if it % 4 else int(rng.integers(1, 1000))
    ih = int(rng.integers(1, 700))
    mode = it % 3
    if mode == 0:      # mild downscale (<= 8 taps)
        ow, oh = max(1, int(iw * rng.uniform(0.5, 1.0))), max(1, int(ih * rng.uniform(0.5, 1.0)))
    elif mode == 1:    # upscale / mixed
        ow, oh = max(1, int(iw * rng.uniform(0.6, 3.0))), max(1, int(ih * rng.uniform(0.4, 3.0)))
    else:              # one axis untouched
        ow, oh = (iw, max(1, int(ih * rng.uniform(0.5, 1.5)))) if it % 2 else (max(1, int(iw * rng.uniform(0.5, 1.5))), ih)
    kind = ["photo", "noise", "noisea", "alpha", "holes"][it % 5]
    img = synth.frame_np(seed * 1000 + it, iw, ih, "photo" if kind == "holes" else kind)
    if kind == "holes":
        for _ in range(5):
            x0, y0 = int(rng.integers(0, iw)), int(rng.integers(0, ih))
            img[y0:y0 + int(rng.integers(1, 80)), x0:x0 + int(rng.integers(1, 80)), 3] = int(rng.choice([0, 0, 1, 254]))
    fmt = it % 2
    ctx.profile(True)
    got = ctx.scale(img, ow, oh, fmt)
    for k in ctx.profile_report():
        kernels[k] = kernels.get(k, 0) + 1
    ctx.profile(False)
    want = oracle.stb_resize(img, ow, oh, fmt)
    if not (got == want).all():
        bad += 1
        print("SCALE MISMATCH", iw, ih, ow, oh, kind, fmt, int(np.abs(got.astype(int) - want).max()), int((got != want).sum()))
print("scale cases", n_scale, "kernels", kernels, "mismatches", bad)
for it in range(n_sixel):
    w, h = int(rng.integers(1, 400)), int(rng.integers(1, 30)) * 6
    kind = ["photo", "noise", "alpha"][it % 3]
    fb = synth.frame_np(seed * 77 + it, w, h, kind)
    fb[..., 3] = 255
    data = ctx.sixel_encode(fb)
    pal, orig, idx = ctx.sixel_debug(w, h)
    _, det = oracle.sixel_encode(fb, True, mode=1)
    img, _ = oracle.sixel_decode(data)
    pct = ((pal.astype(np.int32) * 100 + 127) // 255) * 255 // 100      # what a terminal shows for the palette's percentages
    ok = (orig == det["origcolors"] and pal.shape == det["palette"].shape and (pal == det["palette"]).all()
          and (idx == det["index"]).all() and img.shape == (h, w, 3) and (img == pct[idx]).all())
    if not ok:
        bad += 1
        print("SIXEL MISMATCH", w, h, kind)
print("sixel cases", n_sixel, "total mismatches", bad)
sys.exit(1 if bad else 0)
