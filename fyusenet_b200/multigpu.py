"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed for the rendezvous.

The reference is a single-GPU, batch-1 engine (README.md:72); scaling out is therefore either
  * replicas      -- StyleNet frames <= 1524x1856 and ResNet-50 batch 1: independent requests per GPU, no collective;
  * batch shards  -- ResNet-50 batch B: contiguous image ranges per rank, weights replicated, one all-gather of the
                     [B/world, 1000] logits at the end (NCCL on GPUs, gloo in the CPU tests);
  * row bands     -- StyleNet 4096x4096: horizontal bands.  Implemented as overlapped bands (every rank recomputes
                     the 60 rows of context its band needs, no exchange); the per-layer halo table for an NVLink
                     exchange is computed here as well.
Only host logic lives here; it never touches the oracle and never falls back to CPU compute.
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of `total` independent units (images, frames) owned by `rank`; sizes differ by at most 1."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Step time of the job = the slowest rank (bench.py contract)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_logits(local: np.ndarray, total: int, device=None) -> np.ndarray:
    """All-gather of the per-rank logits [n_local, classes] into [total, classes] in image order.  Shards may differ
    by one image (shard_range), so every rank pads to the largest shard before the collective."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, np.float32)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local.shape[0] == total
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(e - b for b, e in sizes)
    assert local.shape[0] == sizes[rank][1] - sizes[rank][0], "local logits do not match this rank's shard"
    pad = torch.zeros((nmax, local.shape[1]), dtype=torch.float32, device=device)
    pad[:local.shape[0]] = torch.from_numpy(local).to(pad.device)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return np.concatenate([parts[r][:sizes[r][1] - sizes[r][0]].cpu().numpy() for r in range(world)], axis=0)


# ------------------------------------------------------------------------------------------------
# StyleNet row bands: halo rows every layer needs from its band neighbours, at that layer's INPUT
# resolution, derived from the tap geometry of the reference's shaders (fyusenet/gpu/vanilla/convlayerbase_vanilla.cpp:
# 347-371, fractionalconvlayerNxN_vanilla.cpp:43-51).
# ------------------------------------------------------------------------------------------------

def stylenet_halo_plan(ksize: int = 9):
    """[(layer, input scale divisor, halo rows above, halo rows below)] for a band whose first row is a multiple of 4."""
    m = (ksize - 1) // 2
    nres = 5 if ksize == 9 else 2
    plan = [("conv1", 1, m, m), ("conv2", 1, 1, 0), ("conv3", 2, 1, 0)]
    for r in range(1, nres + 1):
        plan += [(f"res{r}_1", 4, 1, 1), (f"res{r}_2", 4, 1, 1)]
    # fractional convs: source row of tap t for output row o is floor(s*(ds*o + 0.5 + t)), t = ky - m
    # deconv3 (s = 0.5, ds = 1): output rows 2j and 2j+1 read source rows floor(j + 0.25 + t/2) and floor(j + 0.75 + t/2)
    above, below = int(np.ceil(m / 2 - 0.25)), int(np.floor(0.75 + m / 2))
    plan += [("deconv1", 4, 1, 0), ("deconv2", 4, 1, 0), ("deconv3", 2, above, below), ("sigmoid", 1, 0, 0)]
    return plan


def band_rows(height: int, world: int, align: int = 4) -> list[tuple[int, int]]:
    """Full-resolution row range [begin, end) per rank; every interior boundary is a multiple of `align` so that the /2
    and /4 levels and the 2x / 4x fractional upsamplers stay on integer rows."""
    if height % align:
        raise ValueError(f"height {height} is not a multiple of {align}")
    units = height // align
    out = []
    for r in range(world):
        b, e = shard_range(units, r, world)
        out.append((b * align, e * align))
    return out


# ------------------------------------------------------------------------------------------------
# StyleNet row bands without an exchange: overlapped bands ("recompute halo")
# ------------------------------------------------------------------------------------------------

def stylenet_margin(ksize: int = 9) -> int:
    """Full-resolution rows of context above / below a band that make every output row of the band exact.

    Walking the halo plan backwards from the output (deconv3 reads +-2 rows at /2, deconv2 and deconv1 one row above
    at /4, ten residual convs +-1 at /4, conv3 / conv2 2o-1..2o+1, conv1 +-m) gives 59 rows above and 51 below for the
    9x9 network (43 / 35 for 3x3); rounded up to a multiple of 4 so that the /2 and /4 levels stay aligned."""
    m = (ksize - 1) // 2
    nres = 5 if ksize == 9 else 2
    d3_above = int(np.ceil(m / 2 - 0.25))            # deconv3, at /2
    d3_below = int(np.floor(0.75 + m / 2))
    # /4 level: deconv2 (floor(o/2) - 1 .. floor(o/2)), deconv1 (o - 1 .. o), 2 * nres residual convs (+-1)
    a4 = (d3_above + 1) // 2 + 1 + 1 + 2 * nres
    b4 = (d3_below + 1) // 2 + 2 * nres
    above = 2 * (2 * a4 + 1) + 1 + m                 # conv3 and conv2: 2o - 1; conv1: m
    below = 2 * (2 * b4 + 1) + 1 + m
    return ((max(above, below) + 3) // 4) * 4


def stylenet_band_plan(height: int, world: int, ksize: int = 9, margin: int | None = None):
    """Per rank: (input row begin, input row end, first output row to keep, rows to keep).  Every rank runs the whole
    network on its band plus `margin` rows of context on each side (clipped at the true image border, where the
    reference's clamp-to-edge applies) and keeps the rows of its own band; no data crosses GPUs."""
    if margin is None:
        margin = stylenet_margin(ksize)
    plan = []
    for b, e in band_rows(height, world):
        ib, ie = max(0, b - margin), min(height, e + margin)
        plan.append((ib, ie, b - ib, e - b))
    return plan


# ------------------------------------------------------------------------------------------------
# StyleNet row bands WITH the per-layer halo exchange over NVLink (fyn_halo_exchange, include/fyusenet_b200.h)
# ------------------------------------------------------------------------------------------------

HALO_MARGIN = 8     # full-resolution rows per band side: 8 / 4 / 2 rows at the /1, /2, /4 levels (an exchange after every layer)
# 44 rows = 11 rows at the /4 level: they cover the ten 3x3 layers of the residual trunk, so the engine (Engine::planHalo) keeps the
# trunk as ONE chain kernel and refreshes the margins twice per frame (behind conv3 and behind res5_2) instead of fifteen times
HALO_MARGIN_SPARSE = 44


def stylenet_halo_margin(ksize: int = 9) -> int:
    """Smallest margin (full-resolution rows, multiple of 4) that covers every layer's taps at that layer's own resolution
    (stylenet_halo_plan): conv1 needs (k-1)/2 rows at /1, deconv3 up to 2 rows at /2, everything else 1 row at its level."""
    need = 4
    for _, div, above, below in stylenet_halo_plan(ksize):
        need = max(need, div * max(above, below))
    return ((need + 3) // 4) * 4


def stylenet_halo_band_plan(height: int, world: int, margin: int = HALO_MARGIN):
    """Per rank: (input row begin, input row end, first output row to keep, rows to keep).  Every rank runs the network on
    its band plus `margin` rows towards each existing neighbour; after every layer the margin rows are replaced by the
    neighbours' band-edge rows, so `margin` only has to cover ONE layer's taps (8 rows instead of the 60 of the
    overlapped-band plan)."""
    plan = []
    for r, (b, e) in enumerate(band_rows(height, world)):
        ib = b - (margin if r > 0 else 0)
        ie = e + (margin if r < world - 1 else 0)
        plan.append((ib, ie, b - ib, e - b))
    return plan


def make_comm(ctx, rank: int | None = None, world: int | None = None):
    """capi.Comm for this process: rank 0 draws the NCCL unique id, torch.distributed (any backend) carries the 128 bytes."""
    import torch.distributed as dist
    from . import capi
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    uid = None
    if world > 1:
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    return capi.Comm(ctx, rank, world, uid)


def bind_to_gpu_numa_node(device_index: int):
    """Pin the calling process to the CPUs that are local to GPU `device_index` (its PCIe root / NUMA node), so that the pinned
    upload / download buffers it allocates afterwards live in that node's memory.  With one process per GPU the end-to-end
    path (two PCIe copies of 79 MB per StyleNet frame) otherwise crosses the socket interconnect for half of the ranks.
    Returns (numa node, number of CPUs) or None when the topology cannot be read; never raises."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        node = int(open(f"{base}/numa_node").read().strip())
        return node, len(cpus)
    except Exception:
        return None
