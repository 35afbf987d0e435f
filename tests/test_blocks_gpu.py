"""GPU parity: CUDA block encoder (through the C ABI) vs the CPU oracle and the golden
fixtures generated from the reference.  Byte-identical or fail."""
import os

import numpy as np
import pytest

import cases
import oracle
import timg_b200
from timg_b200 import synth
from timg_b200.canvas import B200BlockCanvas

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden_blocks():
    return np.load(os.path.join(G, "blocks.npz"))


@pytest.mark.parametrize("name,case", cases.block_cases(), ids=[n for n, _ in cases.block_cases()])
def test_blocks_cuda_matches_golden_and_oracle(ctx, name, case, golden_blocks):
    got = cases.run_block_case(lambda q, u, c: B200BlockCanvas(ctx, q, u, c), case)
    want = cases.run_block_case(lambda q, u, c: oracle.BlockCanvas(q, u, c), case)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == golden_blocks[f"{name}/{i}"].tobytes(), f"{name} frame {i} vs golden"
        assert g == w, f"{name} frame {i} vs oracle"


@pytest.mark.parametrize("seed", range(12))
def test_blocks_cuda_random_sequences(ctx, seed):
    rng = np.random.default_rng(500 + seed)
    q, up, c8 = int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2))
    w = int(rng.integers(1, 400)) * (2 if q else 1)
    h = int(rng.integers(1, 120))
    kind = ["noisea", "photo", "alpha", "noise"][seed % 4]
    frames = [synth.frame_np(2000 + seed, w, h, kind)]
    for k in range(3):
        f = frames[-1].copy()
        n = int(rng.integers(0, 12))
        ys, xs = rng.integers(0, h, n), rng.integers(0, w, n)
        f[ys, xs] = rng.integers(0, 256, (n, 4), dtype=np.uint8)
        frames.append(f)
    case = dict(frames=frames, quarter=q, upper=up, color8=c8, x=int(rng.integers(0, 9)), dy=-h)
    got = cases.run_block_case(lambda *f: B200BlockCanvas(ctx, *f), case)
    want = cases.run_block_case(lambda *f: oracle.BlockCanvas(*f), case)
    assert got == want


def test_blocks_wide_rows_cross_chunk_state(ctx):
    """Rows longer than one 256-thread chunk: SGR state must carry across chunk borders."""
    fb = np.zeros((4, 1500, 4), np.uint8)
    fb[..., 3] = 255
    fb[1, :, 0] = 200                         # one fg colour across the whole row: emitted once
    fb[3, 255:258, 1] = 99                    # change right at a chunk border
    got = B200BlockCanvas(ctx).send(fb)
    assert got == oracle.BlockCanvas().send(fb)


def test_blocks_large_frame_properties(ctx):
    """1080p quarter frame (C3's source size, unscaled): compare with oracle in full."""
    fb = synth.frame_np(77, 1920, 1080, "photo")
    got = B200BlockCanvas(ctx, True).send(fb)
    want = oracle.BlockCanvas(True).send(fb)
    assert len(got) == len(want)
    assert got == want


def test_blocks_identical_frame_is_empty(ctx):
    fb = synth.frame_np(5, 64, 32, "photo")
    cv = B200BlockCanvas(ctx, True)
    assert len(cv.send(fb)) > 0
    assert cv.send(fb, 0, -32) == b""


def test_quarter_odd_width_is_rejected(ctx):
    fb = synth.frame_np(5, 33, 8, "photo")
    with pytest.raises(timg_b200.B200Error) as e:
        ctx.blocks_encode(fb, flags=timg_b200.QUARTER)
    assert e.value.code == timg_b200.EINVAL


def test_small_buffer_reports_needed_size(ctx):
    import ctypes as C
    fb = synth.frame_np(5, 16, 8, "noise")
    buf = C.create_string_buffer(10)
    n = C.c_size_t()
    rc = timg_b200.lib().b200timg_blocks_encode(ctx.h, fb.ctypes.data_as(timg_b200.u8p), 16, 8, None, 0, 0, buf, 10,
                                                C.byref(n))
    assert rc == timg_b200.ENOSPC and n.value == len(oracle.BlockCanvas().send(fb))
