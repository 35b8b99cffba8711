"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def make_scene(sb, ob, n, seed, sh_fmt=0, cov_fmt=0, **kw):
    g = sb.scenes.synthetic_gaussians(n, seed, **kw)
    pods = sb.pack_gaussians(g, sh_fmt, cov_fmt)
    assert np.array_equal(pods, ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE), sh_fmt, cov_fmt))
    return g, pods


def run_frame(sb, ob, ctx, pods, n, cam_args, w, h, sh_fmt=0, cov_fmt=0, mode=0, target_format=0, sh_deg=3,
              no_sh0=False, std_dev=3.0, size=1.0, strict=True, model_transform=None, selection=None, invert=1):
    torch = _torch()
    pos, yaw, pitch = cam_args
    cam = sb.camera_pod(pos, yaw, pitch, w, h)
    ocam = ob.camera_pod(pos, yaw, pitch, w, h)
    assert bytes(cam) == bytes(ocam)
    v = sb.Viewer(ctx, pods, n, sh_fmt=sh_fmt, cov_fmt=cov_fmt, target_format=target_format)
    v.update_camera_with_pod(cam)
    v.update_gaussian_transform(size, mode, sh_deg, no_sh0, std_dev)
    v.set_strict_exp(strict)
    omt = ob.model_transform_pod()
    if model_transform is not None:
        v.update_model_transform(*model_transform)
        omt = ob.model_transform_pod(*model_transform)
    if selection is not None:
        v.enable_selection(True)
        v.set_selection(selection)
        v.set_invert_selection(bool(invert))
    bpp = {0: (torch.uint8, 4), 1: (torch.uint8, 4), 2: (torch.float16, 4), 3: (torch.float32, 4)}[target_format]
    target = torch.zeros((h, w, 4), dtype=bpp[0], device="cuda")
    v.render(target, w, h)
    torch.cuda.synchronize()
    draw, disp = v.read_indirect_args()
    V = int(draw[1])
    idx = v.read_indices(V)
    keys = v.read_depth_keys(sb.padded_key_count(n))
    stats = v.read_frame_stats()
    img = target.cpu().numpy()
    v.close()

    om = ob.OracleModel(pods, n, sh_fmt, cov_fmt, model_transform=omt, selection=selection, invert_selection=invert)
    ogt = ob.gaussian_transform_pod(size, mode, sh_deg, no_sh0, std_dev)
    pre = ob.preprocess(om, ocam, ogt)
    oimg, ostats = ob.render(om, ocam, ogt, target_format, strict_exp=strict)
    return dict(draw=draw, disp=disp, V=V, idx=idx, keys=keys, img=img, stats=stats, pre=pre, oimg=oimg, ostats=ostats)


def check_artifacts(ob, r, n):
    pre = r["pre"]
    assert r["V"] == pre["count"]
    assert list(r["draw"]) == list(pre["draw_args"])
    assert list(r["disp"]) == list(pre["sort_args"])
    V = r["V"]
    # sorted order: oracle = stable LSD sort of the index-ordered compaction
    ok, oi = ob.radix_sort(pre["keys"][:V].view(np.uint32), pre["indices"][:V])
    gk = r["keys"][:V].view(np.uint32)
    assert np.array_equal(gk, ok), "sorted depth keys differ"
    assert np.array_equal(r["idx"], oi), "sorted index order differs (natively index-stable expected)"
    # pad keys (preprocess.wesl:118-125)
    padded = int(pre["sort_args"][0]) * 3840
    assert np.all(r["keys"][V:padded] == np.float32(2.0))
    # visible mask bit-exact
    mask = np.zeros((n + 31) // 32, dtype=np.uint32)
    np.bitwise_or.at(mask, r["idx"] >> 5, np.uint32(1) << (r["idx"] & 31).astype(np.uint32))
    assert np.array_equal(mask, pre["mask"])


@pytest.mark.parametrize("camera", ["outside", "inside"])
def test_frame_small_strict(sb, ob, ctx, camera):
    n = 20000
    g, pods = make_scene(sb, ob, n, 1)
    cam = sb.scenes.CAMERA_OUTSIDE if camera == "outside" else sb.scenes.CAMERA_INSIDE
    r = run_frame(sb, ob, ctx, pods, n, cam, 640, 360)
    check_artifacts(ob, r, n)
    assert not r["stats"]["overflowed"]
    diff = np.abs(r["img"].astype(np.int32) - r["oimg"].astype(np.int32))
    assert diff.max() == 0, f"strict-exp framebuffer must be bit-exact, max diff {diff.max()} at {np.argwhere(diff == diff.max())[:4]}"
    assert np.all(r["img"][..., 3] == 255)


def test_frame_small_fast_exp(sb, ob, ctx):
    n = 20000
    g, pods = make_scene(sb, ob, n, 2)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, 640, 360, strict=False)
    check_artifacts(ob, r, n)
    diff = np.abs(r["img"].astype(np.int32) - r["oimg"].astype(np.int32))
    assert diff.max() <= 2, f"max-abs {diff.max()} > 2/255"
