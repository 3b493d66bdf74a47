#!/usr/bin/env python
"""ResNet-50 batch sharding across the GPUs of one node, run by hand under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tests/mgpu_resnet_dp.py --images 16 --steps 5

One process per GPU: rank r runs images shard_range(B, r, world) through its own engine (weights replicated), the
[B/world, 1000] logits are all-gathered over NCCL (fyusenet_b200.multigpu.gather_logits), rank 0 checks a few images
against the CPU oracle (tolerance of tests/test_gpu_networks.py) and prints the aggregate img/s, timed on the device
and reduced with MAX over ranks.  The gloo/CPU counterpart of the collective logic is tests/test_multigpu_cpu.py.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=2, help="images verified against the oracle on rank 0")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import fyn_oracle as fo
    from fyusenet_b200 import capi, hostapi, multigpu

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b, e = multigpu.shard_range(args.images, rank, world)
    weights = fo.resnet50_synthetic_weights()
    imgs = np.stack([fo.synthetic_image(224, 224, 100 + i) for i in range(b, e)])
    net = hostapi.ResNet50(device=local, batch=e - b)
    net.load_weights(weights)
    net.setup()
    net.set_input(imgs)
    net.forward()                                              # warm-up + result
    local_logits = net.logits().copy()
    full = multigpu.gather_logits(local_logits, args.images, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.forward()                                          # upload + 70 layers + download, synchronous API
    torch.cuda.synchronize()
    ms = multigpu.max_over_ranks((time.perf_counter() - t0) * 1e3, device=dev)
    net.destroy()
    if rank == 0:
        assert full.shape == (args.images, 1000)
        worst = 0.0
        for i in list(range(min(args.check, args.images))) + ([args.images - 1] if args.images > args.check else []):
            ref = fo.resnet50_forward(weights, fo.synthetic_image(224, 224, 100 + i), prec=fo.FP16_STORE)
            err = float(np.linalg.norm(full[i] - ref) / np.linalg.norm(ref))
            assert err <= 5e-3, f"image {i}: rel-L2 {err:.2e}"
            assert set(np.argsort(-full[i])[:5]) == set(np.argsort(-ref)[:5])
            worst = max(worst, err)
        print(json.dumps({"workload": "ResNet-50 224x224 batch-sharded", "images": args.images, "n_gpus": world,
                          "images_per_rank": e - b, "img_per_s": args.images * args.steps / (ms / 1e3), "ms_per_step": ms / args.steps,
                          "worst_rel_l2_vs_oracle": worst, "collective": "NCCL all_gather of [B/world, 1000] logits" if world > 1 else "none"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
