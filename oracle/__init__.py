"""TEST INFRASTRUCTURE ONLY.

ctypes doors onto (a) oracle/liboracle.so -- our CPU restatement of the reference's
hot-path arithmetic (oracle/*.c) and (b) oracle/_ref/libtimg_ref.so -- the reference's
own translation units compiled in place by oracle/Makefile (present only where it was
built: this container, or shipped prebuilt to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (timg_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORC = None
_REF = None


def build(quiet=True):
    """Compile liboracle.so and (if /root/reference is present) _ref/libtimg_ref.so."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _sig(lib, name, res, args):
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = args
    return f


u8p = C.POINTER(C.c_uint8)


def _ptr(a):
    return a.ctypes.data_as(u8p)


def lib():
    global _ORC
    if _ORC is None:
        p = os.path.join(_HERE, "liboracle.so")
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
        if (not os.path.exists(p)) or any(os.path.getmtime(s) > os.path.getmtime(p) for s in srcs):
            build()
        L = C.CDLL(p)
        _sig(L, "orc_as256", C.c_int, [C.c_uint32])
        _sig(L, "orc_compose_bg", None, [u8p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                         C.c_int, C.c_int, C.c_int])
        _sig(L, "orc_calc_fit", C.c_int, [C.c_int] * 6 + [C.c_float] + [C.c_int] * 5 +
             [C.POINTER(C.c_int)] * 2)
        _sig(L, "orc_blocks_new", C.c_void_p, [C.c_int] * 3)
        _sig(L, "orc_blocks_free", None, [C.c_void_p])
        _sig(L, "orc_blocks_bound", C.c_long, [C.c_int, C.c_int])
        _sig(L, "orc_blocks_send", C.c_long, [C.c_void_p, C.c_int, C.c_int, u8p, C.c_int, C.c_int,
                                              C.c_char_p, C.c_int, C.c_char_p])
        _sig(L, "orc_stb_resize", C.c_int, [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int,
                                            C.POINTER(C.c_int)])
        _sig(L, "orc_stb_plan", C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p])
        _ORC = L
    return _ORC


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libtimg_ref.so"))


def ref():
    """The reference itself (UNMODIFIED timg TUs). Raises if it was not built."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libtimg_ref.so")
        if not os.path.exists(p):
            if os.path.exists("/root/reference/src/framebuffer.cc"):
                build()
            else:
                raise RuntimeError("oracle/_ref/libtimg_ref.so not built and /root/reference absent")
        L = C.CDLL(p)
        _sig(L, "ref_calc_fit", C.c_int, [C.c_int] * 6 + [C.c_float] + [C.c_int] * 5 +
             [C.POINTER(C.c_int)] * 2)
        _sig(L, "ref_scale", C.c_int, [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int])
        _sig(L, "ref_compose", None, [u8p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                      C.c_int, C.c_int, C.c_int])
        _sig(L, "ref_as256", C.c_int, [C.c_uint32])
        _sig(L, "ref_parse_color", C.c_uint32, [C.c_char_p])
        _sig(L, "ref_blocks_new", C.c_void_p, [C.c_int] * 4)
        _sig(L, "ref_blocks_prefix", None, [C.c_void_p, C.c_char_p, C.c_int])
        _sig(L, "ref_blocks_send", C.c_long, [C.c_void_p, C.c_int, C.c_int, u8p, C.c_int, C.c_int,
                                              C.c_int, C.c_char_p, C.c_long])
        _sig(L, "ref_blocks_free", None, [C.c_void_p])
        _sig(L, "ref_fb_new", C.c_void_p, [C.c_int, C.c_int])
        _sig(L, "ref_fb_free", None, [C.c_void_p])
        _sig(L, "ref_fb_data", u8p, [C.c_void_p])
        _sig(L, "ref_scale_fb", C.c_int, [C.c_void_p, C.c_void_p])
        _sig(L, "ref_compose_fb", None, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32,
                                         C.c_int, C.c_int, C.c_int])
        _sig(L, "ref_blocks_send_fb", None, [C.c_void_p, C.c_int, C.c_int, C.c_void_p])
        _sig(L, "ref_blocks_flush", None, [C.c_void_p])
        _REF = L
    return _REF


def rgba_u32(r, g, b, a=255):
    return (r & 255) | ((g & 255) << 8) | ((b & 255) << 16) | ((a & 255) << 24)


# ----------------------------------------------------------------- restatement (ours)
def as256(rgba):
    return lib().orc_as256(rgba)


def compose_bg(fb, bg, pattern=0, pw=0, ph=0, start_row=0, has_bg=True):
    out = np.ascontiguousarray(fb, dtype=np.uint8).copy()
    h, w = out.shape[:2]
    lib().orc_compose_bg(_ptr(out), w, h, int(has_bg), bg, pattern, pw, ph, start_row)
    return out


def calc_fit(iw, ih, width, height, cell_x=1, cell_y=2, stretch=1.0, upscale=False,
             upscale_integer=False, fill_width=False, fill_height=False, rotated=False,
             impl=None):
    tw, th = C.c_int(), C.c_int()
    f = (impl or lib().orc_calc_fit)
    r = f(iw, ih, width, height, cell_x, cell_y, stretch, int(upscale), int(upscale_integer),
          int(fill_width), int(fill_height), int(rotated), C.byref(tw), C.byref(th))
    return bool(r), tw.value, th.value


def stb_resize(img, ow, oh, fmt=0, want_info=False):
    """Restatement of ImageScaler::Scale (STB build). info = [vertical_first, h_widest, v_widest, channels]."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    ih, iw = img.shape[:2]
    out = np.empty((oh, ow, 4), np.uint8)
    info = (C.c_int * 4)()
    rc = lib().orc_stb_resize(_ptr(img), iw, ih, fmt, _ptr(out), ow, oh, info)
    assert rc == 0, rc
    return (out, list(info)) if want_info else out


def stb_plan(iw, ih, ow, oh, axis):
    n = ow if axis == 0 else oh
    widest, flags = C.c_int(), C.c_int()
    assert lib().orc_stb_plan(iw, ih, ow, oh, axis, C.byref(widest), C.byref(flags), None, None, None, None) == 0
    first, count, lead = (np.zeros(n, np.int32) for _ in range(3))
    coeff = np.zeros(n * widest.value, np.float32)
    assert lib().orc_stb_plan(iw, ih, ow, oh, axis, None, None, first.ctypes.data, count.ctypes.data,
                              lead.ctypes.data, coeff.ctypes.data) == 0
    return dict(widest=widest.value, flags=flags.value, first=first, count=count, lead=lead,
                coeff=coeff.reshape(n, widest.value))


class BlockCanvas:
    """Restatement of UnicodeBlockCanvas (stateful: backing store, last height/indent)."""

    def __init__(self, quarter=False, upper=False, color8=False):
        self._h = lib().orc_blocks_new(int(quarter), int(upper), int(color8))

    def send(self, fb, x=0, dy=0, prefix=b""):
        fb = np.ascontiguousarray(fb, dtype=np.uint8)
        h, w = fb.shape[:2]
        if dy < 0:
            # Send(): MoveCursorDY(cell_height_for_pixels(dy)) (src/unicode-block-canvas.cc:329,
            # .h:42-45: (pixels - 1) / 2 with C truncation) -> "ESC[nA" (src/terminal-canvas.cc:66-73)
            rows = int((dy - 1) / 2)
            if rows != 0:
                prefix = prefix + (b"\033[%dA" % -rows if rows < 0 else b"\033[%dB" % rows)
        buf = C.create_string_buffer(lib().orc_blocks_bound(w, h) + len(prefix))
        n = lib().orc_blocks_send(self._h, x, dy, _ptr(fb), w, h, prefix, len(prefix), buf)
        return buf.raw[:n]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_blocks_free(self._h)
            self._h = None


# ----------------------------------------------------------------- the reference itself
class RefBlockCanvas:
    """The reference's UnicodeBlockCanvas behind its own BufferedWriteSequencer."""

    def __init__(self, quarter=False, upper=False, color8=False):
        self._h = ref().ref_blocks_new(int(quarter), int(upper), int(color8), 1)

    def prefix(self, data):
        ref().ref_blocks_prefix(self._h, data, len(data))

    def send(self, fb, x=0, dy=0):
        fb = np.ascontiguousarray(fb, dtype=np.uint8)
        h, w = fb.shape[:2]
        cap = 1 << 16
        cap += ((h + 1) // 2) * (16 + w * 39 + 5)
        buf = C.create_string_buffer(cap)
        n = ref().ref_blocks_send(self._h, x, dy, _ptr(fb), w, h, 1, buf, cap)
        assert n >= 0, n
        return buf.raw[:n]

    def __del__(self):
        if getattr(self, "_h", None):
            ref().ref_blocks_free(self._h)
            self._h = None


def ref_scale(img, ow, oh, fmt=0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    ih, iw = img.shape[:2]
    out = np.empty((oh, ow, 4), np.uint8)
    rc = ref().ref_scale(_ptr(img), iw, ih, fmt, _ptr(out), ow, oh)
    assert rc == 0
    return out


def ref_compose_bg(fb, bg, pattern=0, pw=0, ph=0, start_row=0, has_bg=True):
    out = np.ascontiguousarray(fb, dtype=np.uint8).copy()
    h, w = out.shape[:2]
    ref().ref_compose(_ptr(out), w, h, int(has_bg), bg, pattern, pw, ph, start_row)
    return out
