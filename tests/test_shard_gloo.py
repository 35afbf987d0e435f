"""The N>1 path on CPU: two gloo ranks shard frames and gather 'encoded' bytes to rank 0."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from timg_b200 import shard


def test_shard_ranges_cover_and_are_contiguous():
    for n in (0, 1, 7, 8, 64, 1024, 10000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_frames(lo, hi):
    """Deterministic variable-length 'encoded frames' for frame ids lo..hi-1."""
    blobs = []
    for f in range(lo, hi):
        rng = np.random.default_rng(f)
        blobs.append(rng.integers(0, 256, 10 + (f * 37) % 101, dtype=np.uint8).tobytes())
    return blobs


def _worker(rank, world, port, n_frames, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(n_frames, rank, world)
        blobs = _fake_frames(lo, hi)
        payload = torch.frombuffer(bytearray(b"".join(blobs)) or bytearray(1), dtype=torch.uint8)
        offsets = torch.tensor(np.concatenate([[0], np.cumsum([len(b) for b in blobs])]), dtype=torch.int64)
        all_bytes, all_offs = shard.gather_encoded(payload, offsets, dst=0)
        if rank == 0:
            want = _fake_frames(0, n_frames)
            got = [bytes(all_bytes[int(all_offs[i]): int(all_offs[i + 1])].numpy().tobytes()) for i in range(n_frames)]
            ok[0] = int(got == want and all_offs.numel() == n_frames + 1)
        else:
            ok[rank] = int(all_bytes is None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [9, 2, 1])
def test_two_rank_gather_preserves_frame_order(n_frames):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.Array("i", [0, 0])
    procs = [mp.Process(target=_worker, args=(r, 2, port, n_frames, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1, 1]


def _worker_async(rank, world, port, ok):
    """Two gathers in flight (double-buffered payloads), waited in order: what bench.py does so that
    the exchange of batch k overlaps the kernels of batch k+1."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        handles, wants = [], []
        for batch, n_frames in enumerate((7, 4, 5)):
            lo, hi = shard.shard_range(n_frames, rank, world)
            blobs = _fake_frames(100 * batch + lo, 100 * batch + hi)
            payload = torch.frombuffer(bytearray(b"".join(blobs)) or bytearray(1), dtype=torch.uint8)
            offsets = torch.tensor(np.concatenate([[0], np.cumsum([len(b) for b in blobs])]), dtype=torch.int64)
            if len(handles) == 2:
                res = handles.pop(0).wait()
                _check(res, wants.pop(0), rank, ok)
            handles.append(shard.gather_encoded_async(payload, offsets, dst=0))
            wants.append(_fake_frames(100 * batch, 100 * batch + n_frames))
        while handles:
            _check(handles.pop(0).wait(), wants.pop(0), rank, ok)
    finally:
        dist.destroy_process_group()


def _check(res, want, rank, ok):
    all_bytes, all_offs = res
    if rank == 0:
        got = [bytes(all_bytes[int(all_offs[i]): int(all_offs[i + 1])].numpy().tobytes()) for i in range(len(want))]
        ok[0] += int(got == want and all_offs.numel() == len(want) + 1)
    else:
        ok[rank] += int(all_bytes is None)


def test_two_rank_async_gathers_in_flight():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.Array("i", [0, 0])
    procs = [mp.Process(target=_worker_async, args=(r, 2, port, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [3, 3]


def test_animation_chunks_cover_the_sequence_with_one_halo_frame():
    """Delta animations: every rank but the first loads exactly one halo frame and the chunks tile [0, n)."""
    from timg_b200 import shard
    for n in (1, 2, 7, 300, 10000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, cnt, anim = shard.animation_chunk(n, r, world)
                lo, hi = shard.shard_range(n, r, world)
                if hi <= lo:
                    assert cnt == 0
                    continue
                assert anim == (1 if lo == 0 else 2) and first == lo - (anim == 2) and cnt == hi - lo + (anim == 2)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
