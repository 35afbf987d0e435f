"""Multi-GPU plumbing for the frame-sharded path: frames are independent units,
each rank encodes a contiguous chunk on its own B200, and the only exchange is the gather of the
encoded byte buffers to rank 0 (NCCL on GPUs; the same code runs over gloo on CPU for tests).

The reference has no counterpart (single process; frames go one Send() at a time to one tty,
src/renderer.cc:55-58); what is mirrored is its ordering contract: rank 0 ends up with the frames'
bytes concatenated in display order, as BufferedWriteSequencer's FIFO would write them
(src/buffered-write-sequencer.cc:70-89).
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous chunk [lo, hi) of frames for `rank`: sizes differ by at most one."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_encoded(payload, offsets, dst=0, group=None):
    """payload: uint8 tensor holding this rank's frames back to back; offsets: int64 tensor [n+1]
    (same device as payload).  Returns on `dst`: (all_bytes uint8 tensor, all_offsets int64 [N+1])
    with every rank's frames in rank order; on other ranks (None, None).

    One size exchange (all_gather of frame counts and byte totals), one offsets gather, one padded
    payload gather: payloads are ~1 MB/frame against 33 MB/frame of input, so the link is idle
    either way and what matters is the number of collectives, not their size."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = payload.device
    n_local = offsets.numel() - 1
    total_local = int(offsets[-1].item()) if n_local >= 0 else 0
    meta = torch.tensor([n_local, total_local], dtype=torch.int64, device=dev)
    metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    counts = [int(m[0].item()) for m in metas]
    totals = [int(m[1].item()) for m in metas]
    max_total = max(max(totals), 1)
    max_count = max(counts)
    pad_payload = torch.zeros(max_total, dtype=torch.uint8, device=dev)
    pad_payload[:total_local] = payload[:total_local]
    pad_offsets = torch.zeros(max_count + 1, dtype=torch.int64, device=dev)
    pad_offsets[: n_local + 1] = offsets[: n_local + 1].to(torch.int64)
    if rank == dst:
        got_p = [torch.empty(max_total, dtype=torch.uint8, device=dev) for _ in range(world)]
        got_o = [torch.empty(max_count + 1, dtype=torch.int64, device=dev) for _ in range(world)]
    else:
        got_p = got_o = None
    dist.gather(pad_payload, got_p, dst=dst, group=group)
    dist.gather(pad_offsets, got_o, dst=dst, group=group)
    if rank != dst:
        return None, None
    all_bytes = torch.cat([got_p[r][: totals[r]] for r in range(world)])
    offs = [torch.zeros(1, dtype=torch.int64, device=dev)]
    base = 0
    for r in range(world):
        offs.append(got_o[r][1: counts[r] + 1] + base)
        base += totals[r]
    return all_bytes, torch.cat(offs)
