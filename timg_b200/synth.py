"""Deterministic synthetic RGBA frames (integer arithmetic only, so the numpy and the
torch/CUDA generators produce identical bytes).  Harness code for tests and bench.py.

kinds:
  noise  uniform random bytes in R,G,B, A=255            (worst case for encoders)
  noisea uniform random bytes in all four channels
  photo  smooth value-noise (bilinear blend of a coarse random lattice) + +-8 grain, A=255
  alpha  photo colours with a 40-px checker of A in {128, 255}
"""
import numpy as np

_M32 = 0xFFFFFFFF


def _mix_np(x):
    x = x.astype(np.uint64)
    x = (x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D) & np.uint64(_M32)
    x = (x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B) & np.uint64(_M32)
    x = x ^ (x >> np.uint64(16))
    return x


def _hash_np(seed, a, b, c):
    k = (np.uint64(seed & _M32) * np.uint64(0x9E3779B1)) & np.uint64(_M32)
    v = _mix_np(a.astype(np.uint64) + k)
    v = _mix_np((v ^ (b.astype(np.uint64) * np.uint64(0x85EBCA6B))) & np.uint64(_M32))
    v = _mix_np((v ^ (np.uint64(c) * np.uint64(0xC2B2AE35))) & np.uint64(_M32))
    return v


def frame_np(seed, w, h, kind="noise"):
    y, x = np.mgrid[0:h, 0:w]
    out = np.empty((h, w, 4), np.uint8)
    if kind in ("noise", "noisea"):
        for c in range(4):
            out[..., c] = (_hash_np(seed, x, y, c) & np.uint64(255)).astype(np.uint8)
        if kind == "noise":
            out[..., 3] = 255
        return out
    cell = 64
    gx, gy = x // cell, y // cell
    fx, fy = (x % cell).astype(np.int64), (y % cell).astype(np.int64)
    for c in range(3):
        v00 = (_hash_np(seed, gx, gy, 16 + c) & np.uint64(255)).astype(np.int64)
        v10 = (_hash_np(seed, gx + 1, gy, 16 + c) & np.uint64(255)).astype(np.int64)
        v01 = (_hash_np(seed, gx, gy + 1, 16 + c) & np.uint64(255)).astype(np.int64)
        v11 = (_hash_np(seed, gx + 1, gy + 1, 16 + c) & np.uint64(255)).astype(np.int64)
        top = v00 * (cell - fx) + v10 * fx
        bot = v01 * (cell - fx) + v11 * fx
        v = (top * (cell - fy) + bot * fy) // (cell * cell)
        grain = (_hash_np(seed, x, y, 32 + c) & np.uint64(15)).astype(np.int64) - 8
        out[..., c] = np.clip(v + grain, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    if kind == "alpha":
        out[..., 3] = np.where(((x // 40) + (y // 40)) % 2 == 0, 255, 128).astype(np.uint8)
    elif kind != "photo":
        raise ValueError(kind)
    return out


def frames_np(seed, n, w, h, kind="noise"):
    return np.stack([frame_np(seed + i, w, h, kind) for i in range(n)])


# ---------------------------------------------------------------- torch (device) twin
def _mix_t(x):
    x = (x ^ (x >> 16)) * 0x7FEB352D & _M32
    x = (x ^ (x >> 15)) * 0x846CA68B & _M32
    return x ^ (x >> 16)


def _hash_t(seed, a, b, c):
    k = ((seed & _M32) * 0x9E3779B1) & _M32
    v = _mix_t(a + k)
    v = _mix_t((v ^ (b * 0x85EBCA6B)) & _M32)
    v = _mix_t((v ^ (c * 0xC2B2AE35)) & _M32)
    return v


def frame_torch(seed, w, h, kind="noise", device="cuda"):
    import torch
    y = torch.arange(h, device=device, dtype=torch.int64)[:, None].expand(h, w)
    x = torch.arange(w, device=device, dtype=torch.int64)[None, :].expand(h, w)
    out = torch.empty((h, w, 4), dtype=torch.uint8, device=device)
    if kind in ("noise", "noisea"):
        for c in range(4):
            out[..., c] = (_hash_t(seed, x, y, c) & 255).to(torch.uint8)
        if kind == "noise":
            out[..., 3] = 255
        return out
    cell = 64
    gx, gy = x // cell, y // cell
    fx, fy = x % cell, y % cell
    for c in range(3):
        v00 = _hash_t(seed, gx, gy, 16 + c) & 255
        v10 = _hash_t(seed, gx + 1, gy, 16 + c) & 255
        v01 = _hash_t(seed, gx, gy + 1, 16 + c) & 255
        v11 = _hash_t(seed, gx + 1, gy + 1, 16 + c) & 255
        top = v00 * (cell - fx) + v10 * fx
        bot = v01 * (cell - fx) + v11 * fx
        v = (top * (cell - fy) + bot * fy) // (cell * cell)
        grain = (_hash_t(seed, x, y, 32 + c) & 15) - 8
        out[..., c] = torch.clamp(v + grain, 0, 255).to(torch.uint8)
    out[..., 3] = 255
    if kind == "alpha":
        out[..., 3] = torch.where(((x // 40) + (y // 40)) % 2 == 0, 255, 128).to(torch.uint8)
    elif kind != "photo":
        raise ValueError(kind)
    return out
