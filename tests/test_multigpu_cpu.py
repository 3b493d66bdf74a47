"""Host-side multi-process logic on CPU: gloo, world_size 2, rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest

from fyusenet_b200 import multigpu


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 512, 513):
        for world in (1, 2, 3, 8):
            ranges = [multigpu.shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_range(4, 2, 2)


def test_band_rows_and_halo_plan():
    bands = multigpu.band_rows(4096, 8)
    assert bands[0] == (0, 512) and bands[-1] == (3584, 4096)
    assert all(b % 4 == 0 for b, _ in bands)
    bands = multigpu.band_rows(1856, 3)
    assert bands[0][0] == 0 and bands[-1][1] == 1856 and all((e - b) % 4 == 0 for b, e in bands)
    with pytest.raises(ValueError):
        multigpu.band_rows(1855, 2)
    plan = {n: (d, a, b) for n, d, a, b in multigpu.stylenet_halo_plan(9)}
    assert plan["conv1"] == (1, 4, 4)            # 9x9: four rows each side at full resolution
    assert plan["conv2"] == (1, 1, 0)            # 3x3 stride 2 on an even first row: taps 2o-1 .. 2o+1
    assert plan["res3_1"] == (4, 1, 1)
    assert plan["deconv3"] == (2, 2, 2)          # 9 taps spaced half a source texel: rows o/2-2 .. o/2+2
    plan3 = {n: (d, a, b) for n, d, a, b in multigpu.stylenet_halo_plan(3)}
    assert plan3["deconv3"] == (2, 1, 1) and plan3["conv1"] == (1, 1, 1)
    assert len(plan3) == 3 + 4 + 4
    # overlapped bands: 59 rows of context are needed above a band of the 9x9 network (51 below), rounded to 60
    assert multigpu.stylenet_margin(9) == 60 and multigpu.stylenet_margin(3) % 4 == 0
    bp = multigpu.stylenet_band_plan(4096, 4)
    assert bp[0] == (0, 1084, 0, 1024) and bp[1] == (964, 2108, 60, 1024) and bp[3] == (3012, 4096, 60, 1024)


def test_halo_band_plan_covers_every_layer():
    """Row bands with the per-layer exchange: the 8-row margin must cover every layer's taps at that layer's resolution, the
    bands must tile the frame, and every level (/1, /2, /4) must see whole rows."""
    for k in (3, 9):
        assert multigpu.stylenet_halo_margin(k) <= multigpu.HALO_MARGIN
        for _, div, above, below in multigpu.stylenet_halo_plan(k):
            assert max(above, below) <= multigpu.HALO_MARGIN // div
    m = multigpu.HALO_MARGIN
    assert m % 4 == 0
    for height, world in ((4096, 2), (4096, 8), (1856, 3), (64, 2)):
        plan = multigpu.stylenet_halo_band_plan(height, world)
        kept = 0
        for r, (ib, ie, skip, keep) in enumerate(plan):
            assert ib % 4 == 0 and ie % 4 == 0 and keep % 4 == 0
            assert skip == (m if r > 0 else 0) and ie - ib == keep + skip + (m if r < world - 1 else 0)
            assert ib + skip == kept
            kept += keep
        assert kept == height
    assert multigpu.stylenet_halo_band_plan(4096, 8)[3] == (1536 - 8, 2048 + 8, 8, 512)
    # against the overlapped-band plan: 8 rows of margin instead of 60
    assert multigpu.stylenet_margin(9) // multigpu.HALO_MARGIN >= 7


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = multigpu.shard_range(total, rank, world)
        # "logits" of image i: a deterministic row that encodes i
        local = np.stack([np.arange(10, dtype=np.float32) + 100.0 * i for i in range(b, e)]) if e > b else np.zeros((0, 10), np.float32)
        full = multigpu.gather_logits(local, total)
        ms = multigpu.max_over_ranks(10.0 + rank)
        q.put((rank, full, ms))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gather_logits_and_step_time_world2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, total = 2, 7                          # odd total: ranks own 4 and 3 images
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    want = np.stack([np.arange(10, dtype=np.float32) + 100.0 * i for i in range(total)])
    for rank, full, ms in results:
        np.testing.assert_array_equal(full, want)
        assert ms == 11.0                         # max over ranks


def test_halo_band_plan_with_sparse_margin():
    """The 44-row margin of the sparse halo exchange (Engine::planHalo keeps the residual trunk as one chain kernel between two
    exchanges): bands tile the frame, margins exist only towards neighbours, every band height is a multiple of 4 (the /4 level
    needs whole rows) and 11 rows at the /4 level cover the trunk's ten 3x3 layers."""
    m = multigpu.HALO_MARGIN_SPARSE
    assert m % 4 == 0 and m // 4 >= 10 + 1
    for height, world in ((4096, 2), (4096, 4), (4096, 8), (1856, 2)):
        plan = multigpu.stylenet_halo_band_plan(height, world, m)
        assert len(plan) == world
        covered = 0
        for r, (ib, ie, skip, keep) in enumerate(plan):
            assert (ie - ib) % 4 == 0 and keep > 0
            assert skip == (m if r > 0 else 0)
            assert ie - ib == keep + (m if r > 0 else 0) + (m if r < world - 1 else 0)
            assert ib + skip == covered
            covered += keep
        assert covered == height
