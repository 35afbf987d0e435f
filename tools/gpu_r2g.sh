#!/bin/bash
# r2g: parity tests with the emit3 default, A/B of the emitters and of the dither poll interval, ncu of emit3, launch list
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2g
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 emit3 spin256";  $B 2>&1 | tail -1
echo "== C2 emit1 spin256";  B200TIMG_EMIT=1 $B 2>&1 | tail -1
echo "== C2 emit3 spin32";   B200TIMG_DITHER_SPIN=32 $B 2>&1 | tail -1
echo "== C2 emit3 spin128";  B200TIMG_DITHER_SPIN=128 $B 2>&1 | tail -1
echo "== C2 emit3 spin512";  B200TIMG_DITHER_SPIN=512 $B 2>&1 | tail -1
echo "== C2 emit3 spin1000"; B200TIMG_DITHER_SPIN=1000 $B 2>&1 | tail -1
echo "== C2 emit3 parts=1";  B200TIMG_PARTS=1 $B 2>&1 | tail -1
echo "== C5 emit3";          $B --config C5 2>&1 | tail -1
echo "== C4 emit3";          $B --config C4 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resample|compose|sixel|blocks|twopass|yuv" -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sixel_emit3|sixel_dither2" -c 2 -o gpurun_out/${T}_prof -f \
    python bench.py --frames 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
# e2e chunk size of the host pipeline (frames per chunk; default 20 for 4K RGBA)
for c in 8 12 16 20; do echo "== e2e chunk $c"; B200TIMG_CHUNK_FRAMES=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))"; done
