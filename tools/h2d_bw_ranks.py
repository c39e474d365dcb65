"""GPU box, under torchrun: per-rank host->device bandwidth of the 157 MB score tensor from pinned memory when all N
ranks copy at the same time (barrier before every copy) - the ceiling of bench.py's e2e leg at N GPUs.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bw_ranks.py"""
import json, os
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1024 * 9600 * 4
x = torch.empty(n, dtype=torch.float32).pin_memory()
x.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device=f"cuda:{local}")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
gbs = torch.tensor([n * 4 / ts[len(ts) // 2] / 1e6], dtype=torch.float64, device=f"cuda:{local}")
allg = [torch.zeros_like(gbs) for _ in range(dist.get_world_size())]
dist.all_gather(allg, gbs)
if dist.get_rank() == 0:
    v = [round(float(t), 1) for t in allg]
    print(json.dumps({"ranks": dist.get_world_size(), "h2d_GBps_per_rank_concurrent": v, "sum_GBps": round(sum(v), 1),
                      "cpus": os.cpu_count(), "bytes": n * 4}))
dist.barrier()
dist.destroy_process_group()
