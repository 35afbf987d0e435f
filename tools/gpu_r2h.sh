#!/bin/bash
# r2h: tests with TMA staging + identity copy, emit3 timing experiments, ncu of emit3 and of the TMA-staged scaler
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2h
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 emit1 tma";       B200TIMG_EMIT=1 $B 2>&1 | tail -1
echo "== C2 emit1 no tma";    B200TIMG_EMIT=1 B200TIMG_TMA=0 $B 2>&1 | tail -1
echo "== C2 emit3";           $B 2>&1 | tail -1
echo "== C2 emit3 dbg1 (no look-back)";  B200TIMG_E3DBG=1 $B 2>&1 | tail -1
echo "== C2 emit3 dbg2 (no formatting)"; B200TIMG_E3DBG=2 $B 2>&1 | tail -1
echo "== C2 emit3 dbg3";      B200TIMG_E3DBG=3 $B 2>&1 | tail -1
echo "== C5 emit1";           B200TIMG_EMIT=1 $B --config C5 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sixel_emit3" -c 1 -o gpurun_out/${T}_prof_emit3 -f \
    python bench.py --frames 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu emit3 rc=$?"
B200TIMG_EMIT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resample_v3" -c 1 -o gpurun_out/${T}_prof_v3 -f \
    python bench.py --frames 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e >> gpurun_out/${T}_prof.log 2>&1; echo "ncu v3 rc=$?"
