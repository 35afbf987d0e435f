#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kernels-only"
echo "== C2 fast";     $B 2>&1 | tail -1
echo "== C2 exact";    $B --exact-scale 2>&1 | tail -1
echo "== C2 fast dither warps 16"; B200TIMG_DITHER_WARPS=16 $B 2>&1 | tail -1
echo "== C2 fast dither warps 12"; B200TIMG_DITHER_WARPS=12 $B 2>&1 | tail -1
echo "== C2 yuv";      $B --yuv 2>&1 | tail -1
for c in C1 C3 C4 C5; do echo "== $c"; $B --config $c 2>&1 | tail -1; done
timeout 600 python bench.py --steps 5 --warmup 3 --yuv --no-cpu-baseline > gpurun_out/${T}_bench_yuv.json 2> gpurun_out/${T}_bench_yuv.err; echo "yuv bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench_yuv.json").read().strip().splitlines()[-1])
print("yuv value %.0f e2e %.0f pcie bound %.0f h2d GB/s %.1f" % (d["value"], d["e2e"]["value"], d["e2e"]["pcie_bound_mpx_s"], d["e2e"]["pcie_h2d_gbs_measured"]))
PY
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print("C2 value %.0f e2e %.0f cpu %s latency %s" % (d["value"], d["e2e"]["value"], {k: d["cpu_baseline"][k] for k in ("value", "cores", "one_thread")}, d["single_frame_latency"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resample_v3" -c 1 -o gpurun_out/${T}_prof_fast -f \
    python bench.py --frames 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
