"""Per-stage device times (cudaEvents inside the library) for the synthetic configs."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))
import splat_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[1_000_000, 6_000_000])
ap.add_argument("--cams", nargs="+", default=["outside", "inside"])
ap.add_argument("--size", type=int, nargs=2, default=[1920, 1080])
ap.add_argument("--sh", type=int, default=0)
ap.add_argument("--cov", type=int, default=0)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--strip", type=int, nargs=2, default=None, help="row0 rows: render only this strip, with strip culling (one rank of a multi-GPU frame)")
ap.add_argument("--count", action="store_true", help="one extra instrumented frame: fragment / warp-evaluation counters")
args = ap.parse_args()

ctx = sb.Context(0)
w, h = args.size
for n in args.n:
    t0 = time.time()
    g = sb.scenes.synthetic_gaussians(n, sb.scenes.BASE_SEED + 1)
    pods = sb.pack_gaussians(g, args.sh, args.cov)
    del g
    v = sb.Viewer(ctx, pods, n, sh_fmt=args.sh, cov_fmt=args.cov)
    stride = sb.pod_stride(args.sh, args.cov)
    del pods
    print(f"# n={n} built in {time.time() - t0:.1f}s", flush=True)
    v.update_gaussian_transform(1.0, args.mode, 3, False, 3.0)
    v.set_stage_timing(True)
    row0, rows = args.strip if args.strip else (0, 0)
    if args.strip:
        v.set_strip_cull(True)
    target = torch.zeros((rows or h, w, 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    for cam_name in args.cams:
        pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE if cam_name == "outside" else sb.scenes.CAMERA_INSIDE
        v.update_camera(pos, yaw, pitch, w, h)
        acc = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = []
        for it in range(args.iters + 3):
            with torch.cuda.stream(stream):
                e0.record(stream)
                v.render(target, w, h, stream=stream, row0=row0, rows=rows)
                e1.record(stream)
            stream.synchronize()
            if it < 3:
                continue
            tot.append(e0.elapsed_time(e1))
            for k, ms in v.read_stage_times(stream).items():
                acc.setdefault(k, []).append(ms)
        st = v.read_frame_stats(stream)
        V, D = st["visible"], st["duplicates"]
        med = {k: float(np.median(x)) for k, x in acc.items()}
        frame = float(np.median(tot))
        pre_bytes = n * 16 + V * (stride - 16) + 8 * V
        out = dict(path=v.raster_path(), n=n, cam=cam_name, size=[w, h], visible=V, duplicates=D, overflow=st["overflowed"], frame_ms=frame,
                   fps=1000.0 / frame, stages_ms=med,
                   preprocess_GBs=pre_bytes / med["preprocess"] / 1e6, sort_Gkeys=V / med["depth_sort"] / 1e6,
                   sort_GBs=V * 68 / med["depth_sort"] / 1e6, pairs_G=D * 256 / 1e9,
                   raster_Gpairs_s=D * 256 / med["raster"] / 1e6)
        if args.count:
            v.set_raster_counting(True)
            v.render(target, w, h, stream=stream)
            out["counters"] = v.read_raster_counters(stream)
            v.set_raster_counting(False)
        print(json.dumps(out), flush=True)
    v.close()
    del target
