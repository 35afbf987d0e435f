"""TEST INFRASTRUCTURE ONLY.  Real libsixel, if the host happens to have it (SURVEY.md 8c acceptance item 4).

libsixel is the library SixelCanvas::Send calls for the whole quantise / dither / encode step
(src/sixel-canvas.cc:134-148); it is not vendored in the reference and no version is pinned (CMakeLists.txt:44-46).
When `find_library("sixel")` (or pkg-config) finds one, `encode()` runs the reference's exact call sequence

    sixel_output_new(&out, write_cb, buf, NULL);              src/sixel-canvas.cc:135
    sixel_dither_new(&dither, 256, NULL);                     :138
    sixel_dither_initialize(dither, px, w, h, SIXEL_PIXELFORMAT_RGBA8888,
                            SIXEL_LARGE_LUM, SIXEL_REP_AVERAGE_COLORS, SIXEL_QUALITY_AUTO);   :139-142
    sixel_encode(px, w, h, 0, dither, out);                   :144-145

and returns the bytes, so the tests can compare the device (and the restatement) with the real thing.
"""
import ctypes as C
import ctypes.util
import subprocess

_LIB = None
_WHY = None

SIXEL_PIXELFORMAT_RGBA8888 = 0x40 | 0x11       # SIXEL_FORMATTYPE_COLOR | 0x11 (sixel.h)
SIXEL_LARGE_LUM = 0x2
SIXEL_REP_AVERAGE_COLORS = 0x2
SIXEL_QUALITY_AUTO = 0x0


def find():
    """(ctypes library or None, how it was looked for)."""
    global _LIB, _WHY
    if _WHY is None:
        tried = []
        name = ctypes.util.find_library("sixel")
        tried.append(f"find_library('sixel') -> {name}")
        if not name:
            try:
                r = subprocess.run(["pkg-config", "--variable=libdir", "libsixel"], capture_output=True, text=True, timeout=10)
                tried.append(f"pkg-config libsixel -> rc {r.returncode} {r.stdout.strip()}")
                if r.returncode == 0 and r.stdout.strip():
                    name = r.stdout.strip() + "/libsixel.so"
            except Exception as ex:                      # pkg-config itself may be missing
                tried.append(f"pkg-config: {ex}")
        for cand in ([name] if name else []) + ["libsixel.so.1", "libsixel.so"]:
            try:
                _LIB = C.CDLL(cand)
                tried.append(f"loaded {cand}")
                break
            except OSError as ex:
                tried.append(f"{cand}: {ex}")
        _WHY = "; ".join(tried)
    return _LIB, _WHY


def encode(fb):
    """The reference's call sequence on an RGBA8 [h, w, 4] array -> bytes (requires find()[0])."""
    import numpy as np
    L, _ = find()
    fb = np.ascontiguousarray(fb, dtype=np.uint8)
    h, w = fb.shape[:2]
    chunks = []
    WRITE = C.CFUNCTYPE(C.c_int, C.c_char_p, C.c_int, C.c_void_p)

    def cb(data, size, priv):
        chunks.append(C.string_at(data, size))
        return size

    cbf = WRITE(cb)
    out, dither = C.c_void_p(), C.c_void_p()
    L.sixel_output_new.argtypes = [C.POINTER(C.c_void_p), WRITE, C.c_void_p, C.c_void_p]
    L.sixel_dither_new.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    L.sixel_dither_initialize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.sixel_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.sixel_dither_destroy.argtypes = [C.c_void_p]
    L.sixel_output_destroy.argtypes = [C.c_void_p]
    assert L.sixel_output_new(C.byref(out), cbf, None, None) == 0
    assert L.sixel_dither_new(C.byref(dither), 256, None) == 0
    px = fb.copy()                                        # libsixel normalises the pixel format in place
    assert L.sixel_dither_initialize(dither, px.ctypes.data, w, h, SIXEL_PIXELFORMAT_RGBA8888, SIXEL_LARGE_LUM,
                                     SIXEL_REP_AVERAGE_COLORS, SIXEL_QUALITY_AUTO) == 0
    assert L.sixel_encode(px.ctypes.data, w, h, 0, dither, out) == 0
    L.sixel_dither_destroy(dither)
    L.sixel_output_destroy(out)
    return b"".join(chunks)
