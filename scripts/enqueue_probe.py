"""How long does the host take to ENQUEUE a frame (no synchronisation) compared with the device time of the frame?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200")); sys.path.insert(0, ROOT)
import torch
import splat_b200 as sb
import bench
n = 6_000_000
pods = bench.build_scene(n)
ctx = sb.Context(0)
v = sb.Viewer(ctx, pods, n)
v.update_gaussian_transform(1.0, sb.MODE_SPLAT, 3, False, 3.0)
W, H = 1920, 1080
t0 = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda"); t1 = torch.zeros_like(t0)
stream = torch.cuda.Stream()
cams = [bench.view_camera(sb, k) for k in range(64)]
for mode in ("single", "batch"):
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 32
        w0 = time.perf_counter()
        e0.record(stream)
        if mode == "single":
            for i in range(K):
                v.update_camera_with_pod(cams[i]); v.render(t0, W, H, stream=stream)
        else:
            v.render_batch(cams[:K], targets=[(t0, t1)[i & 1] for i in range(K)], width=W, height=H, stream=stream)
        e1.record(stream)
        w1 = time.perf_counter()
        stream.synchronize(); torch.cuda.synchronize()
        w2 = time.perf_counter()
        print(mode, "enqueue ms/frame", round((w1 - w0) * 1e3 / K, 3), "device ms/frame", round(e0.elapsed_time(e1) / K, 3), "wall ms/frame", round((w2 - w0) * 1e3 / K, 3))
