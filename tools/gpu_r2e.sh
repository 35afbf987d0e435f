#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 fast parts=default(2)"; $B 2>&1 | tail -1
echo "== C2 fast parts=1";          B200TIMG_PARTS=1 $B 2>&1 | tail -1
echo "== C2 fast parts=3";          B200TIMG_PARTS=3 $B 2>&1 | tail -1
echo "== C2 fast parts=4";          B200TIMG_PARTS=4 $B 2>&1 | tail -1
echo "== C2 exact";                 $B --exact-scale 2>&1 | tail -1
echo "== C5 parts default";         $B --config C5 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resample_v3|sixel_dither2|sixel_emit_kernel|sixel_palette" -c 4 -o gpurun_out/${T}_prof -f \
    python bench.py --frames 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"yuv420" -c 1 -o gpurun_out/${T}_prof_yuv -f \
    python bench.py --frames 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --yuv > gpurun_out/${T}_prof_yuv.log 2>&1; echo "ncu yuv rc=$?"
