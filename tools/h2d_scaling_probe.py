"""Aggregate host->device ceiling of one node: every rank copies the bench's packed input (48 MB, pinned, allocated after
the rank is bound to its GPU's NUMA node) to its own GPU, all ranks at once.  One JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_scaling_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cpus = bench.pin_to_gpu_numa(bench.physical_gpu_index(local))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 2_400_000 * 20
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 40
e0.record()
for _ in range(iters):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device=dev)
if world > 1:
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    vals = [float(v.item()) for v in allv]
else:
    vals = [gbs]
if rank == 0:
    print(json.dumps({"ranks": world, "h2d_GBps_per_rank": [round(v, 2) for v in vals], "aggregate_GBps": round(sum(vals), 1),
                      "min_GBps": round(min(vals), 2), "numa_cpus_bound": cpus,
                      "frames_per_s_ceiling_at_20B_per_point": round(sum(vals) * 1e9 / (300000 * 20), 0)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
