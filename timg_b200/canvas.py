"""Python mirror of the reference's canvas-level host logic around the C ABI, used by the
tests so they read like calls on the reference classes.  (The C++ adapters a maintainer
would link into timg are in timg_b200/csrc/adapters.h; this file mirrors the same logic.)

UnicodeBlockCanvas::Send host logic restated (src/unicode-block-canvas.cc:323-403):
  * dy < 0 -> queue "ESC[{n}A" with n = -((dy - 1) / 2)            (:329, .h:42-45)
  * x /= 2 in quarter mode                                           (:334)
  * emit_difference = x == last_x && last_h > 0 && |dy| == last_h    (:344-346)
  * if no image byte was produced the whole buffer (prefix included) is dropped (:390-395)
"""
import numpy as np

from . import QUARTER, UPPER, COLOR8


class B200BlockCanvas:
    def __init__(self, ctx, quarter=False, upper=False, color8=False):
        self.ctx = ctx
        self.flags = (QUARTER if quarter else 0) | (UPPER if upper else 0) | (COLOR8 if color8 else 0)
        self.quarter = bool(quarter)
        self.prev = None
        self.last_h = 0
        self.last_x = 0
        self.prefix = b""

    def add_prefix(self, data):
        self.prefix += data

    def send(self, fb, x=0, dy=0):
        fb = np.ascontiguousarray(fb, dtype=np.uint8)
        h, w = fb.shape[:2]
        if dy < 0:
            rows = int((dy - 1) / 2)
            if rows:
                self.prefix += b"\033[%dA" % -rows if rows < 0 else b"\033[%dB" % rows
        prefix, self.prefix = self.prefix, b""
        if self.quarter:
            x //= 2
        emit_diff = (x == self.last_x) and self.last_h > 0 and abs(dy) == self.last_h
        # The backing store equals the previous frame only if the geometry is unchanged.
        prev = self.prev if (emit_diff and self.prev is not None and self.prev.shape == fb.shape) else None
        body = self.ctx.blocks_encode(fb, prev=prev, flags=self.flags, x_indent_cells=x)
        self.prev = fb.copy()
        self.last_h, self.last_x = h, x
        if not body:
            return b""
        return prefix + body
