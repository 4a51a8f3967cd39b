"""Raw PCIe ceiling with N GPUs copying at once (no kernels): the bound `e2e` is held against when N > 1.

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py
Every rank pins 2 x 1 GiB, waits on a barrier, then runs H2D and D2H concurrently on two streams; rank 0 prints one JSON
line with the per-rank and summed rates (GB/s per direction)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


for _ in range(2):
    both()
sync()
reps = 6
t0 = time.perf_counter()
for _ in range(reps):
    both()
torch.cuda.synchronize()
mine = n * reps / (time.perf_counter() - t0) / 1e9
rates = torch.tensor([mine], device="cuda")
if world > 1:
    allr = [torch.zeros_like(rates) for _ in range(world)]
    dist.all_gather(allr, rates)
    allr = [float(x) for x in allr]
else:
    allr = [mine]
if rank == 0:
    print(json.dumps({"probe": "duplex pinned copies, all ranks at once", "n_gpus": world,
                      "per_rank_GBps_each_direction": [round(x, 1) for x in allr],
                      "sum_GBps_each_direction": round(sum(allr), 1),
                      "equivalent_Mpixel_per_s_u16_in_u16_out": round(sum(allr) * 1e9 / 2 / 1e6)}))
if world > 1:
    dist.destroy_process_group()
