#!/bin/bash
# eight GPUs: the default bench line as the driver launches it
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2n8.json 2> gpurun_out/r2n8.err; echo "rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2n8.json').read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('N=8: value %.0f ms %.2f e2e %s' % (d['value'], d['ms_per_step'], e.get('value')))"
tail -3 gpurun_out/r2n8.err
