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
sf = sharding.StripFrame(ctx, v, w, h, 4, world, rank, dst=0, mode=args.mode, balance=True, partition_cull=bool(args.partition_cull))
stream = torch.cuda.current_stream()


def step():
    return sf.render()


for _ in range(args.warmup):
    full = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    full = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    out = dict(config="8K strips", gaussians=args.n, size=[w, h], n_gpus=world, ms_per_frame=float(ms.item()),
               frames_per_s=1000.0 / float(ms.item()), strip_rows=[sharding.strip_rows(h, world, r)[1] for r in range(world)],
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
