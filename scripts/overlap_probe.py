"""Experiment: throughput when two viewers (same scene, own buffers) render alternate views on two streams."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))
import splat_b200 as sb
n, w, h = 6_000_000, 1920, 1080
ctx = sb.Context(0)
pods = sb.pack_gaussians(sb.scenes.synthetic_gaussians(n, sb.scenes.BASE_SEED + 2))
dev = torch.from_numpy(pods).cuda()
res = {}
for nv in (1, 2, 3):
    vs = [sb.Viewer(ctx, n=n, device_pods=dev) for _ in range(nv)]
    streams = [torch.cuda.Stream() for _ in range(nv)]
    targets = [torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda") for _ in range(nv)]
    cams = [sb.camera_pod(*sb.scenes.orbit_camera(k, 64), w, h) for k in range(64)]
    def run(steps):
        for i in range(steps):
            k = i % nv
            vs[k].update_camera_with_pod(cams[i % 64])
            vs[k].render(targets[k], w, h, stream=streams[k])
    run(6); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(60); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[nv] = 60 / dt
    for v in vs: v.close()
print(json.dumps(res))
