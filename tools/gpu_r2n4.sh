#!/bin/bash
# four GPUs: weak scaling of the default bench line, and how many NCCL channels the gather should get
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-4}
R="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline"
show() { python -c "
import json
d=json.loads(open('$1').read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('$2: value %.0f ms %.2f e2e %s' % (d['value'], d['ms_per_step'], e.get('value')))"; }
$R > gpurun_out/r2n${N}_a.json 2> gpurun_out/r2n${N}_a.err; show gpurun_out/r2n${N}_a.json "N=$N nccl defaults"
NCCL_MAX_NCHANNELS=4 $R --no-e2e > gpurun_out/r2n${N}_b.json 2> gpurun_out/r2n${N}_b.err; show gpurun_out/r2n${N}_b.json "N=$N NCCL_MAX_NCHANNELS=4"
NCCL_MAX_NCHANNELS=8 $R --no-e2e > gpurun_out/r2n${N}_c.json 2> gpurun_out/r2n${N}_c.err; show gpurun_out/r2n${N}_c.json "N=$N NCCL_MAX_NCHANNELS=8"
