"""Host -> device bandwidth of this box, one rank per GPU (torchrun): each rank copies a pinned 4 GB buffer to its
GPU, first one rank at a time, then all ranks at once.  Tells whether the end-to-end path (which uploads the score
matrix every call) can scale with the number of GPUs or is capped by the host side of the box."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gb = 4
host = torch.empty(gb << 28, dtype=torch.float32, pin_memory=True)
host.fill_(1.0)
dst = torch.empty_like(host, device=dev)


def bar():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def copy_rate(reps=3):
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return gb * reps / (time.time() - t0)


dst.copy_(host)
alone = None
for r in range(world):
    bar()
    if r == rank:
        alone = copy_rate()
    bar()
bar()
together = copy_rate()
bar()
res = torch.tensor([alone, together], dtype=torch.float64, device=dev)
out = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(out, res)
else:
    out = [res]
if rank == 0:
    print(json.dumps({"h2d_probe": {"ranks": world, "GBps_alone": [round(float(o[0]), 1) for o in out],
                                    "GBps_all_at_once": [round(float(o[1]), 1) for o in out],
                                    "sum_all_at_once": round(sum(float(o[1]) for o in out), 1)}}))
if world > 1:
    dist.destroy_process_group()
