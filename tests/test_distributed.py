"""CPU, world_size 2 over gloo: the host-side sharding logic of the N > 1 paths."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, width, out):
    sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))
    from splat_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the "frame": row index in every channel, so a mis-placed strip is detected
    r0, rows = sharding.strip_rows(height, world, rank)
    strip = torch.arange(r0, r0 + rows, dtype=torch.int32).view(-1, 1, 1).expand(rows, width, 4).contiguous()
    frame = torch.full((height, width, 4), -1, dtype=torch.int32) if rank == 0 else None
    full = sharding.gather_strips(strip, frame, height, world, rank, dst=0)
    # view sharding: every view rendered exactly once; "render" = view id; max-over-ranks timing
    mine = sharding.views_for_rank(64, world, rank)
    counts = torch.zeros(64, dtype=torch.int32)
    counts[mine] += 1
    dist.all_reduce(counts)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        expect = torch.arange(height, dtype=torch.int32).view(-1, 1, 1).expand(height, width, 4)
        out.put((bool(torch.equal(full, expect)), counts.tolist(), float(t.item())))
    else:
        assert full is None
    dist.destroy_process_group()


def test_strip_partition_properties():
    from splat_b200 import sharding
    for height in (1, 15, 16, 17, 360, 1080, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            rows = [sharding.strip_rows(height, world, r) for r in range(world)]
            assert rows[0][0] == 0 and sum(n for _, n in rows) == height
            for (a0, an), (b0, _) in zip(rows, rows[1:]):
                assert a0 + an == b0
            for r0, n in rows:
                assert r0 % 16 == 0 or n == 0


def test_balanced_strips_properties():
    from splat_b200 import sharding
    rng = np.random.default_rng(5)
    for height in (16, 17, 368, 1080, 4320):
        tile_rows = (height + 15) // 16
        for world in (1, 2, 3, 8):
            work = rng.integers(0, 5000, tile_rows) * (rng.random(tile_rows) < 0.7)
            b = sharding.balanced_strips(work, height, world, per_row_cost=3.0)
            assert len(b) == world and b[0][0] == 0 and sum(n for _, n in b) == height
            for (a0, an), (b0, _) in zip(b, b[1:]):
                assert a0 + an == b0 and (b0 % 16 == 0 or b0 == height)

            def load(bounds):
                return max(sum(float(work[y]) + 3.0 for y in range(r0 // 16, (r0 + n + 15) // 16)) if n else 0.0 for r0, n in bounds)

            equal = [sharding.strip_rows(height, world, r) for r in range(world)]
            assert load(b) <= load(equal) + 1e-9, "balanced strips must not be heavier than equal-height strips"


def test_gather_strips_and_view_sharding_gloo_ws2():
    _run_gather(2, 200, 8)


def test_gather_strips_ragged_and_empty_strips_gloo_ws3():
    """17 rows = 2 tile rows over 3 ranks: one rank has no strip at all, one a ragged 1-row strip."""
    _run_gather(3, 17, 4)


def _run_gather(world, height, width):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, counts, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "gathered strips do not reassemble the frame"
    assert counts == [1] * 64, "every view must be rendered by exactly one rank"
    assert tmax == float(world)
