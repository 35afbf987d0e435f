#!/bin/bash
# r2l: tests (incl. the variant-equality tests), the default bench line, kernel tables of every config
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2l
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B="timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
for c in C2 C1 C3 C4 C5; do echo "== $c"; $B --config $c 2>&1 | tail -1; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value %.0f ms %.2f e2e %.0f roof %.4f chain %.4f lat %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['chain']['frac'], d['single_frame_latency']))"
