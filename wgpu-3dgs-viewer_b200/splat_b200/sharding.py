"""Multi-GPU partitioning of the hot path.  One process per GPU.

* view batches (BASELINE config 5a): independent frames, scene replicated, view k -> rank k mod G;
  no data-path collective.
* one very large frame (config 5b): horizontal screen strips, tile-row aligned.  Every rank runs the
  reference's full-frame cull (identical decisions on every rank), keeps only the splats whose tile
  box meets ITS strip (`sb_viewer_set_strip_cull`), depth-sorts, bins and rasterises that subset.
  The strips land in ONE frame on the owner rank:
    - `StripFrame(mode="peer")`: the owner's frame is mapped into every process (CUDA IPC over
      NVLink, `sb_shared_frame_*`) and each rank's rasterizer stores its pixels straight into
      `frame + row0 * pitch` — the kernel's final stores are the transfer; one tiny all-reduce
      orders them before the owner reads.
    - `gather_strips` (mode="sendrecv", also the gloo CPU path): one batched send/recv group that
      receives every strip directly into its rows of the final frame (no padding, no concatenation).
  `StripFrame(partition_cull=True)` (needs the peer mode) also partitions the Preprocessor: rank r culls
  Gaussians [n r / G, n (r + 1) / G) only and stores every strip's splats into the owning rank's
  buffers over NVLink (`sb_strips_*`); one all-reduce separates that scatter from the strips' own
  sort / binning / rasterizer.  Same frames bit for bit.
"""
from __future__ import annotations

TILE = 16


def views_for_rank(n_views: int, world: int, rank: int) -> list[int]:
    return list(range(rank, n_views, world))


def strip_rows(height: int, world: int, rank: int) -> tuple[int, int]:
    """(row0, rows) of rank's strip; boundaries fall on 16-pixel tile rows so no tile is shared.  rows == 0: the rank has no
    strip (more ranks than tile rows) and must not render (SbTarget.rows == 0 would mean the full frame)."""
    tile_rows = (height + TILE - 1) // TILE
    t0 = tile_rows * rank // world
    t1 = tile_rows * (rank + 1) // world
    r0, r1 = min(t0 * TILE, height), min(t1 * TILE, height)
    return r0, r1 - r0


# fixed cost of a tile (CTA start, clear, pixel stores, its share of the tile sort / ranges / schedule) in list entries; measured on
# 8 GPUs with scripts/strips_8k.py --tile-costs (profiles/r02_multigpu.txt)
TILE_COST = 16.0


def balanced_strips(row_work, height: int, world: int, per_row_cost: float = 0.0) -> list[tuple[int, int]]:
    """Strip boundaries (whole tile rows, contiguous, one strip per rank) that minimise the heaviest strip's work, where a tile
    row costs row_work[y] + per_row_cost (row_work = (splat, tile) duplicates per tile row of a calibration / previous frame:
    binning and rasterizer time follow the duplicates, not the rows).  Exact min-max contiguous partition by bisection on the
    bottleneck; deterministic, so every rank that holds the same row_work computes the same strips."""
    tile_rows = (height + TILE - 1) // TILE
    cost = [float(row_work[y]) + per_row_cost for y in range(tile_rows)]
    assert len(cost) == tile_rows

    def parts_needed(limit):
        parts, acc = 1, 0.0
        for c in cost:
            if acc + c > limit and acc > 0.0:
                parts, acc = parts + 1, 0.0
            acc += c
        return parts

    lo, hi = max(cost) if cost else 0.0, sum(cost)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if parts_needed(mid) <= world:
            hi = mid
        else:
            lo = mid
    cuts, acc = [0], 0.0  # greedy fill under the bottleneck `hi`
    for y, c in enumerate(cost):
        if acc + c > hi and acc > 0.0 and len(cuts) < world:
            cuts.append(y)
            acc = 0.0
        acc += c
    while len(cuts) < world:  # fewer strips than ranks: the remaining ranks get empty strips at the end
        cuts.append(tile_rows)
    cuts.append(tile_rows)
    out = []
    for r in range(world):
        r0, r1 = min(cuts[r] * TILE, height), min(cuts[r + 1] * TILE, height)
        out.append((r0, r1 - r0))
    return out


def max_strip_rows(height: int, world: int) -> int:
    return max(strip_rows(height, world, r)[1] for r in range(world))


def gather_strips(strip, frame, height: int, world: int, rank: int, dst: int = 0, bounds=None):
    """Brings every rank's strip (tensor [rows_r, W, C]; may be empty) into `frame` ([H, W, C], on `dst` only) with one
    batched group of point-to-point transfers: the owner receives each strip straight into its rows of the final frame.
    Returns `frame` on dst, None elsewhere.  Works over NCCL (device tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist

    bounds = bounds or [strip_rows(height, world, r) for r in range(world)]
    r0, rows = bounds[rank]
    if world == 1:
        if strip.data_ptr() != frame[r0:r0 + rows].data_ptr():
            frame[r0:r0 + rows].copy_(strip)
        return frame
    ops = []
    if rank == dst:
        if rows and strip.data_ptr() != frame[r0:r0 + rows].data_ptr():
            frame[r0:r0 + rows].copy_(strip)
        for r in range(world):
            if r == dst:
                continue
            q0, qn = bounds[r]
            if qn:
                ops.append(dist.P2POp(dist.irecv, frame[q0:q0 + qn], r))
    elif rows:
        ops.append(dist.P2POp(dist.isend, strip, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return frame if rank == dst else None


class StripFrame:
    """One frame rendered as `world` screen strips into a single buffer on rank `dst` (config 5b)."""

    def __init__(self, ctx, viewer, width: int, height: int, bytes_per_pixel: int, world: int, rank: int, dst: int = 0,
                 mode: str = "peer", balance: bool = False, stream=None, partition_cull: bool = False, tile_cost: float = TILE_COST):
        """balance: rank `dst` renders the full frame once (a calibration frame; a viewer would use its previous frame), reads the
        (splat, tile) duplicates per tile row and broadcasts strips of equal work instead of equal height."""
        import torch
        import torch.distributed as dist
        from . import api

        self.viewer, self.width, self.height, self.world, self.rank, self.dst = viewer, width, height, world, rank, dst
        self.pitch = width * bytes_per_pixel
        self.bpp = bytes_per_pixel
        self.bounds = [strip_rows(height, world, r) for r in range(world)]
        if balance and world > 1:
            box = [None]
            if rank == dst:
                viewer.set_strip_cull(False)
                scratch = torch.zeros((height, width, bytes_per_pixel), dtype=torch.uint8, device="cuda")
                for attempt in range(3):
                    try:
                        viewer.render(scratch, width, height, stream=stream)
                        break
                    except api.SplatError as e:
                        if "render again" not in str(e) or attempt == 2:
                            raise
                tile_rows = (height + TILE - 1) // TILE
                work = viewer.read_tile_row_work(tile_rows, stream)
                del scratch
                # a tile costs its list plus a fixed part (CTA start, clear, store), expressed in list entries (TILE_COST)
                box = [balanced_strips(work, height, world, per_row_cost=tile_cost * ((width + TILE - 1) // TILE))]
            dist.broadcast_object_list(box, src=dst)
            self.bounds = [tuple(b) for b in box[0]]
        self.balanced = bool(balance and world > 1)
        self.row0, self.rows = self.bounds[rank]
        self.mode = mode if world > 1 else "local"
        self.shared = None
        self.strip = None
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        nbytes = height * self.pitch
        if self.mode == "peer":
            try:
                if rank == dst:
                    self.shared = api.SharedFrame(ctx, nbytes)
                    box = [self.shared.handle]
                else:
                    box = [None]
                dist.broadcast_object_list(box, src=dst)
                ok = 1
                if rank != dst:
                    try:
                        self.shared = api.SharedFrame(ctx, handle=box[0])
                    except api.SplatError:
                        ok = 0
                t = torch.tensor([ok], dtype=torch.int32, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                if int(t.item()) == 0:
                    raise api.SplatError("peer mapping unavailable on some rank")
            except api.SplatError:
                if self.shared is not None:
                    self.shared.close()
                    self.shared = None
                self.mode = "sendrecv"
        if self.mode == "peer":
            self.frame_ptr = self.shared.ptr
            self.frame = _wrap(self.shared.ptr, (height, width, bytes_per_pixel)) if rank == dst else None
        else:
            self.frame = torch.zeros((height, width, bytes_per_pixel), dtype=torch.uint8, device="cuda") if rank == dst else None
            if self.mode == "sendrecv" and rank != dst:
                self.strip = torch.zeros((self.rows, width, bytes_per_pixel), dtype=torch.uint8, device="cuda")
        viewer.set_strip_cull(True)
        self.strips = None
        if partition_cull and self.mode == "peer":
            self.strips = api.Strips(viewer, world, rank, self.bounds)
            exports = [None] * world
            dist.all_gather_object(exports, self.strips.exported)
            self.strips.connect(exports)
            dist.barrier()

    def render(self, stream=None):
        """Enqueue this rank's strip and whatever completes the frame on the owner; stream-ordered, no host sync."""
        import torch
        import torch.distributed as dist
        from . import api

        v, w, h = self.viewer, self.width, self.height
        if self.strips is not None:
            ctxm = torch.cuda.stream(stream) if stream is not None else _null()
            for attempt in range(3):
                try:
                    self.strips.scatter(stream)           # K1 on this rank's slice of the model + stores into the peers
                    break
                except api.SplatError as e:
                    if "render again" not in str(e) or attempt == 2:
                        raise
            with ctxm:
                dist.all_reduce(self._flag)               # every rank's scatter has landed
            if self.rows:
                for attempt in range(3):
                    try:
                        self.strips.render(self.frame_ptr + self.row0 * self.pitch, w, h, self.row0, self.rows, self.pitch, stream)
                        break
                    except api.SplatError as e:
                        if "render again" not in str(e) or attempt == 2:
                            raise
            with ctxm:
                dist.all_reduce(self._flag)               # frame complete on the owner; inboxes free for the next scatter
            return self.frame
        if self.rows:
            if self.mode == "sendrecv" and self.rank != self.dst:
                ptr = self.strip.data_ptr()
            else:
                base = self.frame_ptr if self.mode == "peer" else self.frame.data_ptr()
                ptr = base + self.row0 * self.pitch
            for attempt in range(3):
                try:
                    v.render(ptr, w, h, stream=stream, row0=self.row0, rows=self.rows, pitch=self.pitch)
                    break
                except api.SplatError as e:  # an earlier frame overflowed its duplicate buffers: they were grown, render again
                    if "render again" not in str(e) or attempt == 2:
                        raise
        if self.world == 1:
            return self.frame
        ctxm = torch.cuda.stream(stream) if stream is not None else _null()
        with ctxm:
            if self.mode == "peer":
                dist.all_reduce(self._flag)  # stream-ordered fence: every rank's raster stores precede the owner's next read
            else:
                strip = self.strip if self.rank != self.dst else self.frame[self.row0:self.row0 + self.rows]
                gather_strips(strip, self.frame, h, self.world, self.rank, self.dst, self.bounds)
        return self.frame

    def close(self):
        if self.strips is not None:
            self.strips.close()
            self.strips = None
        self.viewer.set_strip_cull(False)
        if self.shared is not None:
            self.shared.close()
            self.shared = None


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _wrap(ptr: int, shape):
    """torch uint8 view over a raw device pointer (the shared frame is cudaMalloc'ed by the library, not by torch)."""
    import torch

    class _Cai:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}

    return torch.as_tensor(_Cai(), device="cuda")
