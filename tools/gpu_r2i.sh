#!/bin/bash
# r2i: tests with the v1b emitter + v3 changes, A/B of the emitters, ncu of both
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2i
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 default (v1b)";   $B 2>&1 | tail -1
echo "== C2 emit v1";         B200TIMG_EMIT=1 $B 2>&1 | tail -1
echo "== C5 default";         $B --config C5 2>&1 | tail -1
echo "== C4 default";         $B --config C4 2>&1 | tail -1
echo "== C2 noise-ish: exact scaler"; $B --exact-scale 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sixel_emit1b|resample_v3" -c 2 -o gpurun_out/${T}_prof -f \
    python bench.py --frames 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
