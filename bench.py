#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   Mpixels/s (INPUT pixels) through  scale -> compose -> median-cut -> FS dither -> sixel emit
workload configs[1]: 3840x2160 RGBA frames -> "-p sixel" on a 300x100-cell terminal (cell 9x18 px,
         src/timg.cc:760-761) -> CalcScaleToFitDisplay -> 2700x1519 -> padded to 1524 rows
         (round_to_sixel, src/sixel-canvas.cc:91-94).  A "step" is one pass of the hot path over one
         batch of --frames distinct synthetic frames (photo-like value noise); the batch is far larger
         than L2 (33 MB/frame), so nothing is cache-resident between steps.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl b200|reference]

Launched under torchrun for N>1 (one rank per GPU): frames are independent units, each rank runs the
same per-GPU batch (weak scaling) and the encoded byte buffers are gathered to rank 0 over NCCL inside
the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IW, IH = 3840, 2160
TERM_COLS, TERM_ROWS, CELL_X, CELL_Y = 300, 100, 9, 18
BG = (0, 0, 0)
KIND = "photo"
SEED = 1234


METRIC = "Mpixels/s scale+dither+sixel-encode @4K\u2192cell"     # BASELINE.json "metric", first clause


def geometry():
    import timg_b200
    _, ow, oh = timg_b200.calc_fit(IW, IH, TERM_COLS * CELL_X, TERM_ROWS * CELL_Y, CELL_X, CELL_Y)
    return ow, oh, (oh + 5) // 6 * 6


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU reference arm
def cpu_frames(n, want_cuda=True):
    """n distinct synthetic frames as numpy arrays (generated on the GPU when there is one: the
    integer generator is identical on both, it is only faster there)."""
    from timg_b200 import synth
    try:
        import torch
        if want_cuda and torch.cuda.is_available():
            return [synth.frame_torch(SEED + i, IW, IH, KIND, "cuda").cpu().numpy() for i in range(n)]
    except Exception:
        pass
    return [synth.frame_np(SEED + i, IW, IH, KIND) for i in range(n)]


def cpu_pipeline_worker(frames, ow, oh, hp, bg_u32, out_sizes, idx):
    """The reference's CPU path for one frame: ImageScaler::Scale (the reference's own STB code,
    oracle/_ref) -> pad + AlphaComposeBackground -> libsixel restatement (oracle mode 0)."""
    import oracle
    for k, fr in frames:
        fb = oracle.ref_scale(fr, ow, oh) if oracle.have_ref() else oracle.stb_resize(fr, ow, oh)
        padded = np.zeros((hp, ow, 4), np.uint8)
        padded[:oh] = fb
        padded = oracle.compose_bg(padded, bg_u32)
        out_sizes[k] = len(oracle.sixel_encode(padded, mode=0))
    idx.append(1)


def run_cpu(frames_per_step, steps, warmup, threads):
    import oracle
    ow, oh, hp = geometry()
    bg = oracle.rgba_u32(*BG)
    oracle.lib()
    pool = cpu_frames(min(4, frames_per_step))
    work = [(k, pool[k % len(pool)]) for k in range(frames_per_step)]
    sizes = {}

    def one_step():
        done = []
        ts = [threading.Thread(target=cpu_pipeline_worker, args=(work[t::threads], ow, oh, hp, bg, sizes, done))
              for t in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    mpx = frames_per_step * steps * IW * IH / 1e6
    return mpx / dt, dt / steps * 1e3, int(np.mean(list(sizes.values())))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=148, help="frames per GPU per step (one batch)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernels-only", action="store_true", help="print just the per-kernel table (tuning runs)")
    ap.add_argument("--exact-scale", action="store_true",
                    help="bit-exact scaler arithmetic instead of the <= 1 LSB fused-multiply-add mode (B200TIMG_FAST_SCALE)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ow, oh, hp = geometry()
    config = {"workload": f"C2: {IW}x{IH} RGBA -> -p sixel, {TERM_COLS}x{TERM_ROWS} cells of {CELL_X}x{CELL_Y}px "
                          f"-> {ow}x{oh} (+pad {hp}) Mitchell scale + compose + 256-colour median cut + FS dither + sixel",
              "frames_per_gpu_per_step": args.frames, "scaled": [ow, oh, hp], "synthetic": KIND,
              "l2": "inputs larger than L2 (33 MB/frame, batch of distinct frames)"}

    if args.impl == "reference":
        # the reference's own CPU implementation of the path on this box's host cores; rank 0 only
        if rank != 0:
            return
        threads = args.cpu_threads or (os.cpu_count() or 1)
        per_step = threads
        value, ms, enc = run_cpu(per_step, max(1, args.steps), max(0, min(args.warmup, 1)), threads)
        import oracle
        kind = "port"   # scaler = the reference's own STB code when oracle/_ref is built; sixel = libsixel restatement
        line = {"impl": "reference", "metric": METRIC, "value": value,
                "unit": "Mpx/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
                "config": dict(config, frames_per_step=per_step),
                "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": threads, "kind": kind,
                                 "sample": f"{per_step} frames/step x {args.steps} steps, one frame per thread; scaler = "
                                           f"{'reference STB TU (oracle/_ref)' if oracle.have_ref() else 'STB restatement'}"
                                           ", sixel = libsixel restatement (libsixel is not vendored)"},
                "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "encoded_bytes_per_frame": enc}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import timg_b200
    from timg_b200 import shard, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(dev)            # a real (non-default) stream shared by torch and the library,
    torch.cuda.set_stream(stream)              # so torch's CUDA events time the library's launches
    ctx = timg_b200.Context(local_rank, stream.cuda_stream)
    assert stream.cuda_stream != 0
    L = timg_b200.lib()
    F = args.frames
    frames = torch.empty((F, IH, IW, 4), dtype=torch.uint8, device=dev)
    for i in range(F):
        frames[i] = synth.frame_torch(SEED + rank * F + i, IW, IH, KIND, dev)
    torch.cuda.synchronize(dev)
    b = timg_b200.Batch(n_frames=F, src_w=IW, src_h=IH, src_fmt=0, out_w=ow, out_h=oh, has_bg=1,
                        bg=timg_b200.rgba_u32(*BG), pattern=0, pattern_w=0, pattern_h=0,
                        flags=0 if args.exact_scale else timg_b200.FAST_SCALE, x_indent_cells=0, animation=0)
    cap = F * 6 * 1024 * 1024
    # two output buffers: with N > 1 the gather of batch k (NCCL, its own stream) runs while batch k+1 is
    # being encoded into the other buffer -- the way a stream of pages / video windows would be served
    nbuf = 2 if world > 1 else 1
    outs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    offss = [torch.zeros(F + 1, dtype=torch.int64, device=dev) for _ in range(nbuf)]
    out, offs = outs[0], offss[0]
    pending = [None] * nbuf
    step_no = [0]

    def drain():
        for i in range(nbuf):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    def step(gather=True):
        i = step_no[0] % nbuf
        step_no[0] += 1
        if pending[i] is not None:              # that buffer's previous batch must have left
            pending[i].wait()
            pending[i] = None
        rc = L.b200timg_sixel_batch_dev(ctx.h, C.byref(b), frames.data_ptr(), outs[i].data_ptr(), cap, offss[i].data_ptr())
        if rc != 0:
            raise RuntimeError(L.b200timg_last_error(ctx.h).decode())
        if world > 1 and gather:
            pending[i] = shard.gather_encoded_async(outs[i], offss[i], dst=0)

    # first call sizes the output; grow the buffers if the guess was too small (write kernel skips, never overruns)
    step(gather=False)
    torch.cuda.synchronize(dev)
    total = int(offss[0][-1].item())
    if total > cap:
        cap = int(total * 1.05)
        outs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        out = outs[0]
    for _ in range(args.warmup):
        step()
    drain()
    torch.cuda.synchronize(dev)
    launches0 = ctx.launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()              # after the sampler start-up, so no rank times another rank's sleep
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    drain()                                    # every batch has arrived on rank 0 inside the timed region
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    total_bytes = int(offss[0][-1].item())
    value = world * F * args.steps * IW * IH / 1e6 / (ms_total / 1e3)

    # ---- per-kernel timing (separate pass, profiling on) -> roofline of the dominant kernel
    roofline, kernels = None, None
    if rank == 0:
        ctx.profile(True)
        for _ in range(2):
            step(gather=False)
        rep = ctx.profile_report()
        ctx.profile(False)
        kernels = {k: {"launches": n, "ms_per_launch": ms / n} for k, (n, ms) in rep.items()}
        chain_ms = sum(ms for _, ms in rep.values()) / 2
        dom = max(rep, key=lambda k: rep[k][1])
        n, ms = rep[dom]
        alg_bytes = F * (4 * IW * IH) + total_bytes          # read every source pixel once + write every encoded byte once
        peak, how = peak_hbm()
        achieved = alg_bytes / (ms / n / 1e3) / 1e9
        traffic = None       # dram__bytes_read+write of that kernel per launch, from the committed ncu capture
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp):
            k = json.load(open(tp))["kernels"].get(dom)
            if k:
                traffic = k["dram_bytes_per_frame"] * F
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": ms / n,
                    "kernel_share_of_chain": (ms / 2) / chain_ms,
                    "chain": {"ms_per_step": chain_ms, "achieved": alg_bytes / (chain_ms / 1e3) / 1e9,
                              "frac": alg_bytes / (chain_ms / 1e3) / 1e9 / peak}}

    # ---- single-frame latency (BASELINE configs[1] is literally one frame): device-resident, batch of 1
    latency = None
    if rank == 0:
        b1 = timg_b200.Batch.from_buffer_copy(b)
        b1.n_frames = 1
        for _ in range(3):
            L.b200timg_sixel_batch_dev(ctx.h, C.byref(b1), frames.data_ptr(), out.data_ptr(), cap, offs.data_ptr())
        torch.cuda.synchronize(dev)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for _ in range(5):
            L.b200timg_sixel_batch_dev(ctx.h, C.byref(b1), frames.data_ptr(), out.data_ptr(), cap, offs.data_ptr())
        l1.record(stream)
        torch.cuda.synchronize(dev)
        latency = {"ms": l0.elapsed_time(l1) / 5, "mpx_s": IW * IH / 1e6 / (l0.elapsed_time(l1) / 5 / 1e3),
                   "note": "one 4K frame through the whole chain; palette + FS wavefront are one CTA per frame"}

    # ---- end to end through the host-buffer ABI call: pinned host frames in, host bytes out
    e2e = None
    if not args.no_e2e:
        Fe = F
        try:
            h_in = torch.empty((Fe, IH, IW, 4), dtype=torch.uint8, pin_memory=True)
        except RuntimeError:
            Fe = max(1, F // 8)
            h_in = torch.empty((Fe, IH, IW, 4), dtype=torch.uint8, pin_memory=True)
        h_in.copy_(frames[:Fe])
        h_out = torch.empty(int(total_bytes * Fe / F * 1.1) + 4096, dtype=torch.uint8, pin_memory=True)
        h_offs = np.zeros(Fe + 1, np.uint64)
        be = timg_b200.Batch.from_buffer_copy(b)
        be.n_frames = Fe

        def e2e_step():
            rc = L.b200timg_sixel_batch(ctx.h, C.byref(be), h_in.data_ptr(), h_out.data_ptr(), h_out.numel(),
                                        h_offs.ctypes.data)
            if rc != 0:
                raise RuntimeError(L.b200timg_last_error(ctx.h).decode())

        e2e_step()
        # raw pinned H2D rate of this box, for context: the e2e number cannot exceed it
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        frames[:min(Fe, 32)].copy_(h_in[:min(Fe, 32)], non_blocking=True)
        torch.cuda.synchronize(dev)
        h2d_gbs = min(Fe, 32) * IW * IH * 4 / (time.perf_counter() - t0) / 1e9
        if world > 1:
            dist.barrier()
        ke = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * Fe * ke * IW * IH / 1e6 / dt, "unit": "Mpx/s",
               "h2d_bytes_per_step": int(Fe * IW * IH * 4), "d2h_bytes_per_step": int(h_offs[Fe]) + 8 * (Fe + 1),
               "frames_per_step": Fe, "steps": ke, "api": "b200timg_sixel_batch (host buffers, pinned)",
               "pcie_h2d_gbs_measured": h2d_gbs,
               "pcie_bound_mpx_s": h2d_gbs * 1e9 / 4 / 1e6}
        del h_in, h_out

    # ---- the reference's CPU path beside it (rank 0, N=1 only): a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or (os.cpu_count() or 1)
        try:
            import oracle
            v, ms_cpu, _ = run_cpu(threads, 1, 0, threads)
            cpu = {"value": v, "unit": "Mpx/s", "cores": threads, "kind": "port",
                   "sample": f"{threads} frames, one per thread, 1 pass ({ms_cpu / 1e3:.1f} s); scaler = "
                             f"{'reference STB TU (oracle/_ref)' if oracle.have_ref() else 'STB restatement'}, "
                             "sixel = libsixel restatement (libsixel is not vendored in the reference)"}
        except Exception as ex:            # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "Mpx/s", "cores": threads, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0 and args.kernels_only:
        print(f"value {value:.0f} Mpx/s  ms/step {ms_total / args.steps:.3f}  " +
              "  ".join(f"{k.replace('sixel_', '').replace('_kernel', '')}={v['ms_per_launch']:.3f}" for k, v in kernels.items()))
    elif rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpx/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
                "config": dict(config, parallelism=f"frames sharded x{world}, NCCL gather of encoded bytes to rank 0; "
                               "double-buffered output: the gather of batch k overlaps the kernels of batch k+1"
                               if world > 1 else "1 GPU"),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "kernels": kernels, "encoded_bytes_per_frame": total_bytes // F,
                "single_frame_latency": latency}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
