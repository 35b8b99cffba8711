"""Stage times for the BASELINE.json configs 2, 2b, 3 and 4 (secondary results; bench.py is the headline)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))
import splat_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", nargs="*", default=None)
args = ap.parse_args()
ctx = sb.Context(0)
stream = torch.cuda.Stream()
FMT = {(0, 0): "single/single 224B", (1, 1): "half/half 128B", (2, 1): "norm8/half 80B"}


def time_viewer(v, w, h, cams, mode):
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.update_gaussian_transform(1.0, mode, 3, False, 3.0)
    v.set_stage_timing(True)
    out = {}
    for name, (pos, yaw, pitch) in cams.items():
        v.update_camera(pos, yaw, pitch, w, h)
        acc, tot = {}, []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(args.iters + 3):
            e0.record(stream)
            v.render(target, w, h, stream=stream)
            e1.record(stream)
            stream.synchronize()
            if it < 3:
                continue
            tot.append(e0.elapsed_time(e1))
            for k, ms in v.read_stage_times(stream).items():
                acc.setdefault(k, []).append(ms)
        st = v.read_frame_stats(stream)
        out[name] = dict(frame_ms=float(np.median(tot)), fps=1000 / float(np.median(tot)), visible=st["visible"],
                         duplicates=st["duplicates"], overflow=st["overflowed"],
                         stages_ms={k: round(float(np.median(x)), 4) for k, x in acc.items()})
    return out


def single(tag, n, seed, sh, cov, w, h, modes=(0,)):
    if args.only and tag not in args.only:
        return
    g = sb.scenes.synthetic_gaussians(n, seed)
    pods = sb.pack_gaussians(g, sh, cov)
    del g
    v = sb.Viewer(ctx, pods, n, sh_fmt=sh, cov_fmt=cov)
    stride = sb.pod_stride(sh, cov)
    for mode in modes:
        r = time_viewer(v, w, h, {"outside": sb.scenes.CAMERA_OUTSIDE, "inside": sb.scenes.CAMERA_INSIDE}, mode)
        for cam, d in r.items():
            V = d["visible"]
            pre = n * 16 + V * (stride - 16) + 8 * V
            d.update(config=tag, n=n, pod=FMT[(sh, cov)], size=[w, h], mode=["splat", "ellipse", "point"][mode], camera=cam,
                     preprocess_GBs=round(pre / d["stages_ms"]["preprocess"] / 1e6, 1),
                     sort_Gkeys_s=round(V / d["stages_ms"]["depth_sort"] / 1e6, 2))
            print(json.dumps(d), flush=True)
    v.close()


single("config2_1M_1080p", 1_000_000, sb.scenes.BASE_SEED + 1, 0, 0, 1920, 1080)
single("config2b_6M_1080p", 6_000_000, sb.scenes.BASE_SEED + 2, 0, 0, 1920, 1080)
single("config3_6M_half_4K", 6_000_000, sb.scenes.BASE_SEED + 3, 1, 1, 3840, 2160, modes=(0, 1))
single("config3_6M_norm8_4K", 6_000_000, sb.scenes.BASE_SEED + 3, 2, 1, 3840, 2160, modes=(0, 1))

if not args.only or "config4_multimodel" in args.only:
    # config 4: 8 models x 1M, per-model transforms, draw order by descending centroid distance and a
    # permuted order, and a selection mask of rectangle density (25 %) with invert = 0 / 1
    w, h, n = 1920, 1080, 1_000_000
    mm = sb.MultiModelViewer(ctx)
    pos, yaw, pitch = (0.0, 0.0, -60.0), 0.0, 0.0
    mm.update_camera_with_pod(sb.camera_pod(pos, yaw, pitch, w, h))
    cents = []
    for k in range(8):
        pods = sb.pack_gaussians(sb.scenes.synthetic_gaussians(n, sb.scenes.BASE_SEED + 4 + k))
        mm.insert_model(k, pods, n)
        a = 0.3 * k
        t = (12.0 * (k - 3.5), 0.0, 0.0)
        mm.update_model_transform_with_pod(k, sb.model_transform_pod(t, (0.0, float(np.sin(a / 2)), 0.0, float(np.cos(a / 2))), (1 + 0.05 * k,) * 3))
        cents.append(np.linalg.norm(np.array(t) - np.array(pos)))
    far_first = [int(i) for i in np.argsort(cents)[::-1]]
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for label, order, sel in (("far-to-near", far_first, None), ("permuted", [3, 0, 7, 1, 6, 2, 5, 4], None),
                              ("far-to-near + mask invert=0", far_first, 0), ("far-to-near + mask invert=1", far_first, 1)):
        if sel is not None:
            rng = np.random.default_rng(9)
            for k in range(8):
                bits = (rng.random((n + 31) // 32 * 32) < 0.25).astype(np.uint8)
                mm.set_selection(k, np.packbits(bits, bitorder="little").view(np.uint32), invert=bool(sel))
        tot = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(args.iters + 3):
            e0.record(stream)
            mm.render(target, w, h, order, stream=stream)
            e1.record(stream)
            stream.synchronize()
            if it >= 3:
                tot.append(e0.elapsed_time(e1))
        print(json.dumps(dict(config="config4_multimodel", models=8, n_each=n, size=[w, h], order=label,
                              frame_ms=float(np.median(tot)), fps=1000 / float(np.median(tot)))), flush=True)
    mm.close()
