"""2-rank probe: send/recv and gather_encoded timing breakdown (prints on rank 0)."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from timg_b200 import shard
r = int(os.environ["RANK"]); w = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 600 << 20
x = torch.empty(n, dtype=torch.uint8, device=dev)
def timeit(fn, iters=4):
    fn(); fn(); torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / iters
def batched():
    ops = [dist.P2POp(dist.isend, x, 0)] if r == 1 else [dist.P2POp(dist.irecv, x, 1)]
    for q in dist.batch_isend_irecv(ops): q.wait()
big = torch.empty(n * w, dtype=torch.uint8, device=dev)
def agt(): dist.all_gather_into_tensor(big, x)
offs = torch.arange(0, 149, dtype=torch.int64, device=dev) * (n // 148)
def ge(): shard.gather_encoded(x, offs, dst=0)
res = []
for name, fn in (("batch_isend_irecv 600MB", batched), ("all_gather_into_tensor 600MB/rank", agt), ("gather_encoded 600MB", ge)):
    t = timeit(fn)
    res.append(f"{name:36s} {t*1e3:8.2f} ms  {n/t/1e9:7.1f} GB/s")
if r == 0:
    print("\n".join(res))
dist.destroy_process_group()
