#!/bin/bash
# Round-2 first GPU pass: parity tests on the new emit/dither kernels, A/B kernel timings, ncu captures.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== default";              $B 2>&1 | tail -1
echo "== EMIT_MATCH=ballot";    B200TIMG_EMIT_MATCH=b $B 2>&1 | tail -1
echo "== EMIT_V1";              B200TIMG_EMIT_V1=1 $B 2>&1 | tail -1
echo "== DITHER_V1";            B200TIMG_DITHER_V1=1 $B 2>&1 | tail -1
echo "== frames 16 default";    timeout 300 python bench.py --frames 16 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sixel_emit2|sixel_dither2|resample_planar" -c 3 -o gpurun_out/${T}_prof -f \
    python bench.py --frames 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${T}_bench.json
