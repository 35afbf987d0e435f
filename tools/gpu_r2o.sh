#!/bin/bash
# single-frame latency against the number of CTAs one frame's FS wavefront is split over
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for s in default 6 12 16 24 48; do
  if [ $s = default ]; then unset B200TIMG_DITHER_SPLIT; else export B200TIMG_DITHER_SPLIT=$s; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split $s: latency %.3f ms  value %.0f' % (d['single_frame_latency']['ms'], d['value']))"
done
for n in 2 4 8 16; do
  unset B200TIMG_DITHER_SPLIT
  timeout 300 python bench.py --frames $n --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only 2>/dev/null | tail -1 | cut -c1-200
done
