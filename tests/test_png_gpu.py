"""GPU tests of the PNG + base64 path behind the kitty / iTerm2 canvases (timg_b200/csrc/png.cu).

Reference: png::Encode, src/timg-png.cc:90-152 (Sub filter, one IDAT, CRC per chunk) + EncodeBase64,
src/timg-base64.h:28-53.  The reference compresses with libdeflate (third party, not in its tree): the compressed
bytes are unpinnable, the decoded PIXELS and the container fields are not.  Parity here = the stream parses with
Python's own zlib / struct, every chunk CRC and the Adler-32 verify, the filter byte of every row is 1 (Sub) as in
the reference, it decodes to exactly the source pixels, and the base64 text equals base64.b64encode.
"""
import base64
import struct
import zlib

import numpy as np
import pytest

import timg_b200
from timg_b200 import synth

pytestmark = pytest.mark.gpu


def png_decode(data):
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) == crc, typ
        chunks.append((typ, body))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]             # the reference writes exactly these (:105-151)
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, comp, filt, inter) == (8, 0, 0, 0) and ctype in (2, 6)
    bpp = 4 if ctype == 6 else 3
    raw = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + w * bpp)   # zlib checks the Adler-32
    assert (raw[:, 0] == 1).all()                                              # kFilterType = 1, "Sub" (:92)
    px = np.cumsum(raw[:, 1:].reshape(h, w, bpp).astype(np.uint32), axis=1).astype(np.uint8)   # undo Sub: prefix sums mod 256
    return px, ctype


@pytest.mark.parametrize("w,h,kind", [(67, 50, "alpha"), (1, 1, "noise"), (320, 90, "photo"), (333, 201, "noisea"),
                                      (2700, 25, "photo"), (16390, 4, "noise")])
def test_png_decodes_to_the_source_pixels(ctx, w, h, kind):
    fb = synth.frame_np(3 + w, w, h, kind)
    for rgb24 in (False, True):
        data, b64 = ctx.png_encode(fb, rgb24)
        assert len(data) == timg_b200.lib().b200timg_png_size(w, h, int(rgb24))
        px, ctype = png_decode(data)
        assert ctype == (2 if rgb24 else 6)
        assert (px == (fb[..., :3] if rgb24 else fb)).all()
        assert b64 == base64.b64encode(data)
