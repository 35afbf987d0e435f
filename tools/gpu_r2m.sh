#!/bin/bash
# r2m: round-end pass -- parity tests, smoke, both bench arms, every config, launch list, full ncu capture of the C2 chain
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2m
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "reference arm rc=$?"
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for c in C1 C3 C4 C5; do $B --config $c > gpurun_out/${T}_bench_$c.json 2>> gpurun_out/${T}_bench.err; echo "== $c rc=$?"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resample|compose|sixel|blocks|twopass|yuv" -c 200 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resample_v3|sixel_dither2|sixel_emit|sixel_palette|sixel_lut|sixel_compact" -c 6 -o gpurun_out/${T}_prof -f \
    python bench.py --frames 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e", d.get("e2e") and round(d["e2e"].get("value")),
              "roof", d.get("roofline") and round(d["roofline"]["frac"], 4), {k: round(v["ms_per_launch"], 3) for k, v in (d.get("kernels") or {}).items()})
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
