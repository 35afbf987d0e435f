#!/bin/bash
# r2k: tests (staged two-pass first pass), v3 build variants, C1/C4 with and without the staged pass
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2k
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 default (sent1 minb3)"; $B 2>&1 | tail -1 | cut -c1-120
for v in sent0_minb2 sent0_minb3 sent1_minb2; do echo "== C2 $v"; B200TIMG_LIBFILE=$PWD/timg_b200/libb200timg_$v.so $B 2>&1 | tail -1 | cut -c1-120; done
echo "== C1 staged"; $B --config C1 2>&1 | tail -1
echo "== C1 plain";  B200TIMG_NO_H1S=1 $B --config C1 2>&1 | tail -1
echo "== C4 staged"; $B --config C4 2>&1 | tail -1
echo "== C4 plain";  B200TIMG_NO_H1S=1 $B --config C4 2>&1 | tail -1
echo "== C2 parts=1"; B200TIMG_PARTS=1 $B 2>&1 | tail -1 | cut -c1-60
echo "== C2 parts=2"; B200TIMG_PARTS=2 $B 2>&1 | tail -1 | cut -c1-60
echo "== C5 parts=1"; B200TIMG_PARTS=1 $B --config C5 2>&1 | tail -1 | cut -c1-60
echo "== C5 parts=4"; B200TIMG_PARTS=4 $B --config C5 2>&1 | tail -1 | cut -c1-60
