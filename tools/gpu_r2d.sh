#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2d
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 fast";     $B 2>&1 | tail -1
echo "== C2 exact";    $B --exact-scale 2>&1 | tail -1
echo "== C2 yuv";      $B --yuv 2>&1 | tail -1
for c in C1 C3 C4 C5; do echo "== $c"; $B --config $c 2>&1 | tail -1; done
echo "== C3 yuv"; $B --config C3 --yuv 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --yuv --no-cpu-baseline > gpurun_out/${T}_bench_yuv.json 2> gpurun_out/${T}_bench_yuv.err; echo "yuv bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2d_bench_yuv.json").read().strip().splitlines()[-1])
print("yuv value %.0f e2e %.0f pcie bound %.0f" % (d["value"], d["e2e"]["value"], d["e2e"]["pcie_bound_mpx_s"]))
PY
echo "== cpu scaling (reference arm, C2)"
for t in 1 8 32 64 128; do timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --cpu-threads $t | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($t, 'threads', round(d['value'],1), 'Mpx/s')"; done
