#!/usr/bin/env python
"""PCIe check for the end-to-end arm: H2D of one input frame (33.9 MB) and D2H of one output frame (45.3 MB), alone and
concurrently on two streams (pinned host memory).

    python tools/pcie_overlap.py                                    one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_overlap.py
                                                                    N ranks, one GPU each, copying AT THE SAME TIME (barrier before
                                                                    every measurement): where does the end-to-end curve flatten --
                                                                    per-GPU link, host memory, or a shared root complex?
Rank 0 prints one JSON line: per-rank milliseconds and the aggregate GB/s of the concurrent pair."""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n_in, n_out = 1524 * 1856 * 3, 1524 * 1856 * 4
h_in = torch.empty(n_in, dtype=torch.float32).pin_memory()
h_out = torch.empty(n_out, dtype=torch.float32).pin_memory()
d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
d_out = torch.empty(n_out, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=30):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
mine = torch.tensor([a, b, c], dtype=torch.float64, device=dev)
if world > 1:
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    allv = [v.tolist() for v in allv]
else:
    allv = [mine.tolist()]
if rank == 0:
    worst = [max(v[i] for v in allv) for i in range(3)]
    print(json.dumps({"ranks": world, "h2d_33.9MB_ms": [round(v[0], 3) for v in allv], "d2h_45.3MB_ms": [round(v[1], 3) for v in allv],
                      "both_ms": [round(v[2], 3) for v in allv],
                      "aggregate_GBps": {"h2d_alone": round(world * 33.94 / worst[0], 1), "d2h_alone": round(world * 45.26 / worst[1], 1),
                                         "both": round(world * (33.94 + 45.26) / worst[2], 1)},
                      "frames_per_s_bound_float_io": round(world * 1e3 / worst[2], 1)}))
if world > 1:
    dist.destroy_process_group()
