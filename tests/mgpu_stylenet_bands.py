#!/usr/bin/env python
"""StyleNet 9x9 on one large frame split into row bands over the GPUs of one node, by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/mgpu_stylenet_bands.py --width 4096 --height 4096 --steps 5

Default: overlapped bands (fyusenet_b200.multigpu.stylenet_band_plan): rank r uploads its band plus 60 rows of context per side
from the host frame, runs the whole network on it and keeps its own rows; nothing crosses GPUs on the data path.
--halo: per-layer halo exchange over NVLink (multigpu.stylenet_halo_band_plan + Engine::setHaloExchange -> fyn_halo_exchange):
8-row margins that the neighbours refresh with peer stores after every layer.
Rank 0 gathers the bands' checksums (NCCL all_gather) and, with --verify, compares its band with the rows of a
whole-frame run on its own GPU (bit-exact).  Throughput = frames / max-over-ranks device time.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--height", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--halo", action="store_true", help="per-layer NVLink halo exchange instead of recomputed context rows")
    ap.add_argument("--kernel", type=int, default=9)
    ap.add_argument("--margin", type=int, default=0, help="halo margin in full-resolution rows (default: multigpu.HALO_MARGIN_SPARSE; 8 = an exchange after every layer)")
    ap.add_argument("--device-resident", action="store_true", help="time the layers + exchanges only (upload / download layers skipped)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from fyusenet_b200 import capi, hostapi, multigpu, synthetic

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K = args.kernel
    weights = synthetic.stylenet_weights(K)
    img = synthetic.image(args.height, args.width, 7)
    margin = args.margin or multigpu.HALO_MARGIN_SPARSE
    plan = multigpu.stylenet_halo_band_plan(args.height, world, margin) if args.halo else multigpu.stylenet_band_plan(args.height, world, K)
    ib, ie, skip, keep = plan[rank]
    net = hostapi.StyleNet(K, args.width, ie - ib, device=local)
    net.load_weights(weights)
    net.setup()
    comm = None
    if args.halo:
        comm = multigpu.make_comm(capi.Context(local), rank, world)
        net.set_halo_exchange(comm, margin, ie - ib)
    exchanges, chained = (net.halo_exchanges if args.halo else 0), net.chained_layers
    net.set_input(img[ib:ie])
    net.forward()
    band = net.output_rgba()[0][skip:skip + keep].copy()
    if args.device_resident:
        net.skip_io(True)
        net.forward()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.forward()                                    # upload of the band + layers + download, synchronous API
    torch.cuda.synchronize()
    ms = multigpu.max_over_ranks((time.perf_counter() - t0) * 1e3, device=dev)
    pushed = comm.info()["halo_bytes_pushed"] // (args.steps + 1) if comm is not None else 0
    net.destroy()
    if comm is not None:
        comm.destroy()
    sums = torch.tensor([float(band[..., :3].astype(np.float64).sum())], dtype=torch.float64, device=dev)
    parts = [torch.zeros_like(sums) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, sums)
    else:
        parts = [sums]
    exact = None
    if args.verify:
        whole = hostapi.StyleNet(K, args.width, args.height, device=local)
        whole.load_weights(weights)
        whole.setup()
        whole.set_input(img)
        whole.forward()
        ref = whole.output_rgba()[0][ib + skip:ib + skip + keep]
        exact = bool(np.array_equal(ref, band))
        whole.destroy()
        flag = torch.tensor([1 if exact else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exact = bool(flag.item())
    if rank == 0:
        mode = "row bands with per-layer NVLink halo exchange" if args.halo else "overlapped row bands"
        print(json.dumps({"workload": f"StyleNet {K}x{K} {args.width}x{args.height}, {world} {mode}", "n_gpus": world,
                          "frames_per_s": args.steps / (ms / 1e3), "ms_per_frame": ms / args.steps, "band_rows": keep,
                          "context_rows": margin if args.halo else multigpu.stylenet_margin(K), "exchanges_per_frame": exchanges, "chained_layers": chained,
                          "device_resident": bool(args.device_resident), "nvlink_bytes_pushed_per_frame_rank0": int(pushed),
                          "rgb_checksum": float(sum(p.item() for p in parts)), "bit_exact_vs_whole_frame": exact}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
