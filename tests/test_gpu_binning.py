"""GPU: the binning kernels (K4' dup_count / dup_offsets, K5' dup_emit2) against the round-1 pair and against the oracle.

The library reads its binning switches (SB_BIN, SB_BIN_TILES_PREFIX) once per process, so every variant renders in its own
subprocess and reports digests of its frames; the default variant's frames are also checked against the CPU oracle here."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (name, n, seed, extent, log_scale, camera, width, height, row0, rows)
CASES = [
    ("small_splats_outside", 200_000, 11, 10.0, (-5.0, -3.0), "outside", 1920, 1080, 0, 0),
    # near-camera splats: boxes of hundreds to thousands of tiles — splats that span many 4096-duplicate emit windows, windows that
    # hold a single splat, the warp-cooperative expansion, and runs of splats with no tile at all (cut away by the alpha cut-off)
    ("huge_splats_inside", 3_000, 12, 2.0, (-2.5, 0.5), "inside", 2560, 1440, 0, 0),
    ("mixed_inside", 150_000, 13, 10.0, (-5.0, -1.5), "inside", 1920, 1080, 0, 0),
    # a strip of the frame: boxes are clipped to the strip's tile rows
    ("strip_rows", 120_000, 14, 10.0, (-4.5, -2.5), "outside", 1920, 1080, 352, 208),
    ("one_gaussian", 1, 15, 1.0, (-3.0, -3.0), "outside", 640, 360, 0, 0),
]

WORKER = r"""
import hashlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(sys.argv[1], "wgpu-3dgs-viewer_b200"))
import splat_b200 as sb
cases = json.loads(sys.argv[2])
ctx = sb.Context(0)
out = {}
for name, n, seed, extent, log_scale, cam, w, h, row0, rows in cases:
    g = sb.scenes.synthetic_gaussians(n, seed, extent=extent, log_scale=tuple(log_scale))
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE if cam == "outside" else sb.scenes.CAMERA_INSIDE
    v = sb.Viewer(ctx, pods, n)
    v.set_strict_exp(True)
    v.update_camera(pos, yaw, pitch, w, h)
    def frame():
        # a frame that needs more duplicates than the viewer reserved is flagged, the next enqueue reports it (status 7) after
        # growing the buffers, and the frame after that is complete
        for attempt in range(5):
            target = torch.zeros((rows or h, w, 4), dtype=torch.uint8, device="cuda")
            try:
                if rows:
                    v.render(target, w, h, row0=row0, rows=rows)
                else:
                    v.render(target, w, h)
            except sb.SplatError as e:
                assert e.status == 7, e
                continue
            torch.cuda.synchronize()
            st = v.read_frame_stats()
            if not st["overflowed"]:
                return target, st
        raise AssertionError("the frame kept overflowing")
    digests = []
    for it in range(2):  # the second frame reuses every buffer of the first
        target, st = frame()
        digests.append(hashlib.sha256(target.cpu().numpy().tobytes()).hexdigest())
    out[name] = dict(digests=digests, visible=st["visible"], duplicates=st["duplicates"])
    v.close()
print("RESULT " + json.dumps(out))
"""


def run_variant(env_extra):
    env = dict(os.environ)
    for k in ("SB_BIN", "SB_BIN_TILES_PREFIX"):
        env.pop(k, None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT, json.dumps(CASES)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def render_kwargs_supported(sb):
    import inspect
    return "row0" in inspect.signature(sb.Viewer.render).parameters


def test_binning_variants_render_identical_frames(sb):
    if not render_kwargs_supported(sb):
        pytest.skip("Viewer.render has no strip arguments")
    default = run_variant({})
    v1 = run_variant({"SB_BIN": "v1"})
    prefix = run_variant({"SB_BIN_TILES_PREFIX": "1"})
    for name, *_ in CASES:
        d = default[name]
        assert d["digests"][0] == d["digests"][1], f"{name}: a second frame differs from the first"
        for other, tag in ((v1, "SB_BIN=v1"), (prefix, "SB_BIN_TILES_PREFIX=1")):
            o = other[name]
            assert o["visible"] == d["visible"] and o["duplicates"] == d["duplicates"], (name, tag, o, d)
            assert o["digests"] == d["digests"], f"{name}: frame under {tag} differs from the default binning"
    # the cases do what they are meant to: boxes far beyond one emit window, and a frame with many more duplicates than splats
    assert default["huge_splats_inside"]["duplicates"] > 40 * default["huge_splats_inside"]["visible"]
    assert default["one_gaussian"]["visible"] == 1


ORACLE_CASES = [
    ("huge_splats_inside_small", 300, 12, 2.0, (-2.5, 0.5), "inside", 960, 544, 0, 0),   # sized for the CPU oracle
    ("strip_rows_small", 30_000, 14, 10.0, (-4.5, -2.5), "outside", 960, 544, 176, 112),
]


@pytest.mark.parametrize("case", ORACLE_CASES, ids=lambda c: c[0])
def test_default_binning_equals_the_oracle(sb, ob, ctx, case):
    """Strict exp: the frame is bit-identical to the CPU oracle's (the strip case against the oracle's rows)."""
    import torch
    if not render_kwargs_supported(sb):
        pytest.skip("Viewer.render has no strip arguments")
    name, n, seed, extent, log_scale, cam, w, h, row0, rows = case
    g = sb.scenes.synthetic_gaussians(n, seed, extent=extent, log_scale=log_scale)
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE if cam == "outside" else sb.scenes.CAMERA_INSIDE
    v = sb.Viewer(ctx, pods, n)
    v.set_strict_exp(True)
    v.update_camera(pos, yaw, pitch, w, h)
    target = torch.zeros((rows or h, w, 4), dtype=torch.uint8, device="cuda")
    if rows:
        v.render(target, w, h, row0=row0, rows=rows)
    else:
        v.render(target, w, h)
    torch.cuda.synchronize()
    img = target.cpu().numpy()
    v.close()
    oimg, _ = ob.render(ob.OracleModel(pods, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
    ref = oimg[row0:row0 + rows] if rows else oimg
    assert np.array_equal(img, ref)
