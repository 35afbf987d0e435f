#!/bin/bash
# e2e throughput of the host-buffer batch call for a few pipeline chunk sizes (tuning aid)
for c in "$@"; do
  echo -n "chunk=$c  "
  B200TIMG_CHUNK_FRAMES=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('value %.0f  e2e %.0f  (pcie bound %.0f)' % (d['value'], d['e2e']['value'], d['e2e']['pcie_bound_mpx_s']))"
done
