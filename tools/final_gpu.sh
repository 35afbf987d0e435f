#!/bin/bash
# Round-end GPU pass (one gpurun call): parity tests, smoke, both bench arms, ncu launch list + full capture.
# Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ afterwards.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${1:-r1}
python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resample|compose|sixel|blocks" -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"resample|sixel" -c 10 -o gpurun_out/${T}_prof -f \
    python bench.py --frames 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_prof.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/${T}_bench.json", "gpurun_out/${T}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f" % d["value"], "e2e", d.get("e2e", {}) and d["e2e"].get("value"), "cpu", d.get("cpu_baseline"))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
