"""Config 5b: ONE 7680x4320 frame split into horizontal screen strips, one per GPU, gathered with
NCCL (torch.distributed) to rank 0.  Launch: torchrun --nproc-per-node G scripts/strips_8k.py
(G = 1 runs the single-GPU frame).  Every rank runs the full-frame cull, then keeps, depth-sorts, bins and
rasterises only the splats of its own strip (splat_b200/sharding.py: StripFrame)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))
import splat_b200 as sb  # noqa: E402
from splat_b200 import sharding  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=6_000_000)
ap.add_argument("--size", type=int, nargs=2, default=[7680, 4320])
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--mode", default="peer", choices=["peer", "sendrecv"])
ap.add_argument("--partition-cull", type=int, default=1, help="1: sb_strips_* (each rank preprocesses a slice of the model), 0: replicated cull")
ap.add_argument("--tile-costs", type=float, nargs="*", default=None,
                help="sweep of the balancer's fixed cost per tile (list entries): per-rank stage times for each")
ap.add_argument("--check", action="store_true", help="rank 0 also renders the full frame alone and compares")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h = args.size
ctx = sb.Context(local)
pods = sb.pack_gaussians(sb.scenes.synthetic_gaussians(args.n, sb.scenes.BASE_SEED + 5))
v = sb.Viewer(ctx, pods, args.n)
pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
v.update_camera(pos, yaw, pitch, w, h)
stream = torch.cuda.current_stream()


def measure(sf):
    for _ in range(args.warmup):
        full = sf.render()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        full = sf.render()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), full


if args.tile_costs:
    # the balancer's cost model: duplicates + tile_cost per tile.  For every candidate: the frame time and every rank's own
    # stage times (cudaEvents inside the library), so the slowest rank and what it spends its time on are visible.
    for tc in args.tile_costs:
        sf = sharding.StripFrame(ctx, v, w, h, 4, world, rank, dst=0, mode=args.mode, balance=True, partition_cull=False, tile_cost=tc)
        ms, _ = measure(sf)
        v.set_stage_timing(True)
        sf.render()
        st = v.read_stage_times(stream)
        v.set_stage_timing(False)
        mine = dict(rank=rank, rows=sf.rows, sum_ms=round(sum(st.values()), 4), **{k: round(x, 4) for k, x in st.items()})
        allr = [None] * world
        if world > 1:
            dist.all_gather_object(allr, mine)
        else:
            allr = [mine]
        if rank == 0:
            print(json.dumps(dict(tile_cost=tc, n_gpus=world, ms_per_frame=ms, per_rank=allr)), flush=True)
        if world > 1:
            dist.barrier()
        sf.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)

sf = sharding.StripFrame(ctx, v, w, h, 4, world, rank, dst=0, mode=args.mode, balance=True, partition_cull=bool(args.partition_cull))
ms_val, full = measure(sf)
if rank == 0:
    out = dict(config="8K strips", gaussians=args.n, size=[w, h], n_gpus=world, ms_per_frame=ms_val,
               frames_per_s=1000.0 / ms_val, strip_rows=[b[1] for b in sf.bounds],
               transport=sf.mode, partitioned_cull=sf.strips is not None)
    if args.check:
        full = full.clone()
        v.set_strip_cull(False)
        ref = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        v.render(ref, w, h, stream=stream)
        torch.cuda.synchronize()
        out["identical_to_single_gpu_frame"] = bool(torch.equal(ref, full))
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
