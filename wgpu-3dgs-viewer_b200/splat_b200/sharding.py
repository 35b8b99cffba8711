"""Multi-GPU partitioning of the hot path.  One process per GPU.

* view batches (BASELINE config 5a): independent frames, scene replicated, view k -> rank k mod G;
  no data-path collective.
* one very large frame (config 5b): horizontal screen strips, tile-row aligned; every rank runs the
  full-frame cull + depth sort (identical artefacts on every rank), bins and rasterises only its
  strip, then ONE gather of the strips over torch.distributed (NCCL over NVLink on GPUs, gloo in
  the CPU tests).
"""
from __future__ import annotations

TILE = 16


def views_for_rank(n_views: int, world: int, rank: int) -> list[int]:
    return list(range(rank, n_views, world))


def strip_rows(height: int, world: int, rank: int) -> tuple[int, int]:
    """(row0, rows) of rank's strip; boundaries fall on 16-pixel tile rows so no tile is shared."""
    tile_rows = (height + TILE - 1) // TILE
    t0 = tile_rows * rank // world
    t1 = tile_rows * (rank + 1) // world
    r0, r1 = min(t0 * TILE, height), min(t1 * TILE, height)
    return r0, r1 - r0


def max_strip_rows(height: int, world: int) -> int:
    return max(strip_rows(height, world, r)[1] for r in range(world))


def gather_strips(strip, height: int, world: int, rank: int, dst: int = 0):
    """Gathers per-rank strips (torch tensors [rows_r, W, C]) into the full frame on `dst`.
    Strips are padded to a common row count so a single fixed-size gather is used."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return strip
    pad_rows = max_strip_rows(height, world)
    buf = torch.zeros((pad_rows,) + tuple(strip.shape[1:]), dtype=strip.dtype, device=strip.device)
    buf[: strip.shape[0]] = strip
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst)
        rows = [strip_rows(height, world, r)[1] for r in range(world)]
        return torch.cat([p[:n] for p, n in zip(parts, rows)], dim=0)
    dist.gather(buf, None, dst=dst)
    return None
