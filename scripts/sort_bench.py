"""Standalone RadixSorter benchmark: Gkeys/s on uniform random f32 depth keys and on u32 keys."""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wgpu-3dgs-viewer_b200"))
import splat_b200 as sb

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[1_000_000, 6_000_000, 64_000_000])
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
ctx = sb.Context(0)
for n in args.n:
    for dist in ("f32_depth", "u32"):
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        if dist == "u32":
            keys = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g) * 2
            bits = 32
        else:
            keys = torch.rand(n, device="cuda", generator=g).view(torch.int32)
            bits = 32
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
        s = sb.RadixSorter(ctx, n)
        k, v = keys.clone(), vals.clone()
        stream = torch.cuda.Stream()
        times = []
        for it in range(args.iters + 3):
            k.copy_(keys); v.copy_(vals)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                s.sort(k.data_ptr(), v.data_ptr(), cnt.data_ptr(), n, 0, bits, stream=stream)
                e1.record(stream)
            stream.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        ok = None
        if args.check:
            ku = keys.cpu().numpy().view(np.uint32)
            order = np.argsort(ku, kind="stable")
            ok = bool(np.array_equal(k.cpu().numpy().view(np.uint32), ku[order]) and np.array_equal(v.cpu().numpy().view(np.uint32), order.astype(np.uint32)))
        ms = float(np.median(times))
        print(json.dumps(dict(n=n, dist=dist, ms=ms, gkeys_s=n / ms / 1e6, gbs=n * 68 / ms / 1e6, frac_hbm=n * 68 / ms / 1e6 / 6541.8, ok=ok)), flush=True)
        s.close()
