#!/bin/bash
# two GPUs: which gather transport keeps up with the 13 ms chain?
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
R="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2: value %.0f ms %.2f' % (d['value'], d['ms_per_step']))"; }
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P $R > gpurun_out/r2n3_a.json 2> gpurun_out/r2n3_a.err; show gpurun_out/r2n3_a.json "abi gather, CUDA_MEMCPY=1 (default)"
grep -E "via P2P|via SHM|via NET|NVLS|CUDA_MEMCPY|P2P/" gpurun_out/r2n3_a.err | head -8
NCCL_P2P_USE_CUDA_MEMCPY=0 $R > gpurun_out/r2n3_b.json 2> gpurun_out/r2n3_b.err; show gpurun_out/r2n3_b.json "abi gather, CUDA_MEMCPY=0"
$R --py-gather > gpurun_out/r2n3_c.json 2> gpurun_out/r2n3_c.err; show gpurun_out/r2n3_c.json "py gather (torch p2p), CUDA_MEMCPY=1"
NCCL_P2P_USE_CUDA_MEMCPY=0 $R --py-gather > gpurun_out/r2n3_d.json 2> gpurun_out/r2n3_d.err; show gpurun_out/r2n3_d.json "py gather (torch p2p), CUDA_MEMCPY=0"
