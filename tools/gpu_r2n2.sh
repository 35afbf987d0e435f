#!/bin/bash
# two GPUs of one box: the multi-rank bench line (NCCL gather through the C ABI), N=1 right before for the ratio
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_1gpu.err; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2n_bench_2gpu.json 2> gpurun_out/r2n_2gpu.err; echo "n2 rc=$?"; tail -3 gpurun_out/r2n_2gpu.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/r2n_bench_{n}gpu.json").read().strip().splitlines()[-1])
        print(n, "gpus: value %.0f ms %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["config"].get("parallelism"))
    except Exception as ex:
        print(n, "unreadable", ex)
PY
