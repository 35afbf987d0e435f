#!/bin/bash
# four GPUs with the high-priority gather stream
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=4
R="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline"
show() { python -c "
import json
d=json.loads(open('$1').read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('$2: value %.0f ms %.2f e2e %s' % (d['value'], d['ms_per_step'], e.get('value')))"; }
$R --no-e2e > gpurun_out/r2n4p_a.json 2> gpurun_out/r2n4p_a.err; show gpurun_out/r2n4p_a.json "N=4 priority stream, nccl defaults"
NCCL_MAX_NCHANNELS=8 $R --no-e2e > gpurun_out/r2n4p_b.json 2> gpurun_out/r2n4p_b.err; show gpurun_out/r2n4p_b.json "N=4 priority stream, NCCL_MAX_NCHANNELS=8"
