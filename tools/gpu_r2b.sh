#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2b
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== fast scale";   $B 2>&1 | tail -1
echo "== exact scale";  $B --exact-scale 2>&1 | tail -1
echo "== NO_V3 exact";  B200TIMG_NO_V3=1 $B --exact-scale 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resample_v3" -c 1 -o gpurun_out/${T}_prof_fast -f \
    python bench.py --frames 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resample_v3" -c 1 -o gpurun_out/${T}_prof_exact -f \
    python bench.py --frames 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --exact-scale > gpurun_out/${T}_prof2.log 2>&1; echo "ncu rc=$?"
