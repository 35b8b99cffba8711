"""GPU parity at BASELINE.json's own configs, at their full sizes: the CUDA path (through the C ABI) against the CPU oracle.

config 1   coverage/model.ply at 1280x720, examples/simple.rs defaults — loaded with sb_read_ply (and the golden pins)
config 2b  6 M Gaussians, pod single/single, 1920x1080, splat (the headline scene): artefacts + framebuffer bit-exact
config 3   6 M Gaussians, pods half/half and norm8/half, 3840x2160, splat and ellipse
config 4   8 models x 1 M, per-model transforms, far-to-near and permuted order, rectangle mask from sb_mm_select_rect, invert 0/1
config 5b  one 7680x4320 frame: a screen strip against the oracle's render of the same rows
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

_SCENES = {}


def scene(sb, n, seed):
    """Synthetic source Gaussians, generated once per (n, seed) for the whole module (6 M take ~10 s of numpy)."""
    if (n, seed) not in _SCENES:
        if len(_SCENES) >= 2:  # keep at most two big scenes (1.3 GB each at 6 M) alive
            _SCENES.pop(next(iter(_SCENES)))
        _SCENES[(n, seed)] = sb.scenes.synthetic_gaussians(n, seed)
    return _SCENES[(n, seed)]


def idiff(a, b):
    return int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max())


def read_artifacts(sb, v, n):
    draw, disp = v.read_indirect_args()
    V = int(draw[1])
    return dict(draw=draw, disp=disp, V=V, idx=v.read_indices(V), keys=v.read_depth_keys(sb.padded_key_count(n)))


def check_artifacts(ob, art, pre, n):
    """visible mask, indirect counts, keys, pads and the sorted index order: bit-exact (north star)."""
    V = art["V"]
    assert V == pre["count"]
    assert list(art["draw"]) == list(pre["draw_args"]) and list(art["disp"]) == list(pre["sort_args"])
    ok, oi = ob.radix_sort(pre["keys"][:V].view(np.uint32), pre["indices"][:V])
    assert np.array_equal(art["keys"][:V].view(np.uint32), ok), "sorted depth keys differ"
    assert np.array_equal(art["idx"], oi), "sorted index order differs"
    assert np.all(art["keys"][V:int(pre["sort_args"][0]) * 3840] == np.float32(2.0))
    mask = np.zeros((n + 31) // 32, dtype=np.uint32)
    np.bitwise_or.at(mask, art["idx"] >> 5, np.uint32(1) << (art["idx"] & 31).astype(np.uint32))
    assert np.array_equal(mask, pre["mask"]), "visible mask differs"


# ------------------------------------------------------------------------------------------ config 1

def test_config1_model_ply_720p(sb, ob, ctx, tmp_path):
    """coverage/model.ply (the reference's only model fixture; its 9 x 62 vertex properties are committed under
    tests/golden/) through sb_read_ply -> sb_viewer_create_from_gaussians -> sb_viewer_render with the
    examples/simple.rs:163-186 defaults: model rotated 180 degrees about Z, camera at the origin, 1280x720."""
    import torch
    gold = json.load(open(os.path.join(HERE, "golden", "golden.json")))
    props = np.load(os.path.join(HERE, "golden", "model_ply_props.npy"))
    names = (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)]
             + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(props)
    hdr += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    path = tmp_path / "model.ply"
    path.write_bytes(hdr.encode() + props.astype("<f4").tobytes())
    g = sb.read_ply(str(path))
    assert len(g) == 9 and g.tobytes() == ob.gaussians_from_ply_props(props).tobytes()

    w, h = 1280, 720
    half = np.float32(np.pi) / 2
    rot = (0.0, 0.0, float(np.sin(half)), float(np.cos(half)))
    om = ob.OracleModel(ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE)), 9, model_transform=ob.model_transform_pod((0, 0, 0), rot, (1, 1, 1)))
    ocam, ogt = ob.camera_pod((0, 0, 0), 0.0, 0.0, w, h), ob.gaussian_transform_pod()
    pre = ob.preprocess(om, ocam, ogt)
    pins = gold["oracle"]["model_ply"]
    for strict in (True, False):
        v = sb.Viewer(ctx, gaussians=g)
        v.update_camera((0, 0, 0), 0.0, 0.0, w, h)
        v.update_model_transform((0, 0, 0), rot, (1, 1, 1))
        v.set_strict_exp(strict)
        target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        v.render(target, w, h)
        art = read_artifacts(sb, v, 9)
        check_artifacts(ob, art, pre, 9)
        # SURVEY 8(c) known answers and the committed pins
        assert sorted(art["idx"].tolist()) == gold["survey_kat"]["model_ply"]["visible"] == [1, 2, 3, 5, 6, 8]
        assert art["idx"].tolist() == pins["indices_sorted"]
        assert [f"0x{int(k):08x}" for k in art["keys"][:art["V"]].view(np.uint32)] == pins["keys_sorted"]
        oimg, ost = ob.render(om, ocam, ogt, strict_exp=strict)
        img = target.cpu().numpy()
        d = idiff(img, oimg)
        assert d == 0 if strict else d <= 2, f"strict={strict}: max-abs {d}/255"
        if not strict:  # the pins were generated with the oracle's libm exp
            assert [int(oimg[..., c].astype(np.int64).sum()) for c in range(4)] == pins["image_sum_rgba"]
        assert img[..., :3].max() > 0 and np.all(img[..., 3] == 255)
        v.close()


# ------------------------------------------------------------------------------------------ config 2b

def test_config2b_6M_1080p_bit_exact(sb, ob, ctx):
    """The headline scene itself: visible set, counts, keys, order and (reproducible exp) the framebuffer bit-exact against
    the oracle; the default MUFU path within 2/255."""
    import torch
    n, w, h = 6_000_000, 1920, 1080
    g = scene(sb, n, sb.scenes.BASE_SEED + 2)
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    om = ob.OracleModel(pods, n)
    ocam, ogt = ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod()
    threads = ob.use_all_host_threads()
    pre = ob.preprocess(om, ocam, ogt)
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for strict in (True, False):
        v.set_strict_exp(strict)
        v.render(target, w, h)
        torch.cuda.synchronize()
        if strict:
            check_artifacts(ob, read_artifacts(sb, v, n), pre, n)
            assert not v.read_frame_stats()["overflowed"]
        oimg, _ = ob.render(om, ocam, ogt, strict_exp=strict, n_threads=threads)
        d = idiff(target.cpu().numpy(), oimg)
        assert d == 0 if strict else d <= 2, f"strict={strict}: max-abs {d}/255"
    v.close()


# ------------------------------------------------------------------------------------------ config 3

@pytest.mark.parametrize("sh_fmt,cov_fmt", [(1, 1), (2, 1)])
def test_config3_6M_compressed_4K(sb, ob, ctx, sh_fmt, cov_fmt):
    """6 M Gaussians in the compressed pods (half/half 128 B, norm8/half 80 B) at 3840x2160, splat and ellipse modes."""
    import torch
    n, w, h = 6_000_000, 3840, 2160
    g = scene(sb, n, sb.scenes.BASE_SEED + 3)
    pods = sb.pack_gaussians(g, sh_fmt, cov_fmt)
    assert np.array_equal(pods[: 4096 * sb.pod_stride(sh_fmt, cov_fmt)],
                          ob.pack_gaussians(g[:4096].view(ob.GAUSSIAN_DTYPE), sh_fmt, cov_fmt))
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    om = ob.OracleModel(pods, n, sh_fmt, cov_fmt)
    ocam = ob.camera_pod(pos, yaw, pitch, w, h)
    threads = ob.use_all_host_threads()
    v = sb.Viewer(ctx, pods, n, sh_fmt=sh_fmt, cov_fmt=cov_fmt)
    v.update_camera(pos, yaw, pitch, w, h)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for mode in (sb.MODE_SPLAT, sb.MODE_ELLIPSE):
        ogt = ob.gaussian_transform_pod(display_mode=mode)
        v.update_gaussian_transform(1.0, mode, 3, False, 3.0)
        v.render(target, w, h)
        torch.cuda.synchronize()
        if mode == sb.MODE_SPLAT:
            check_artifacts(ob, read_artifacts(sb, v, n), ob.preprocess(om, ocam, ogt), n)
        assert not v.read_frame_stats()["overflowed"]
        oimg, _ = ob.render(om, ocam, ogt, n_threads=threads)
        d = idiff(target.cpu().numpy(), oimg)
        assert d <= 2, f"mode {mode}: max-abs {d}/255"
    v.close()


# ------------------------------------------------------------------------------------------ config 4

def test_config4_multi_model_8x1M_rect_mask(sb, ob, ctx):
    """8 models x 1 M with the SURVEY 8(d) transforms at 1080p: far-to-near (examples/multi_model.rs:254-270) and a permuted
    draw order, then the rectangle [480,1440)x[270,810) evaluated on the GPU for every model (sb_mm_select_rect) with
    invert = 0 (only the selected Gaussians are shown) and invert = 1 (they are hidden)."""
    import torch
    n, w, h = 1_000_000, 1920, 1080
    pos, yaw, pitch = sb.scenes.CAMERA_CONFIG4
    cam, ocam = sb.camera_pod(pos, yaw, pitch, w, h), ob.camera_pod(pos, yaw, pitch, w, h)
    ogt = ob.gaussian_transform_pod()
    threads = ob.use_all_host_threads()
    mm = sb.MultiModelViewer(ctx)
    mm.update_camera_with_pod(cam)
    pods_of, mts = {}, {}
    for k in range(8):
        pods_of[k] = sb.pack_gaussians(sb.scenes.synthetic_gaussians(n, sb.scenes.BASE_SEED + 4 + k))
        assert mm.insert_model(k, pods_of[k], n) is False
        mts[k] = sb.scenes.config4_transform(k)
        mm.update_model_transform(k, *mts[k])

    def omodels(order, sel=None, invert=1):
        return [ob.OracleModel(pods_of[k], n, model_transform=ob.model_transform_pod(*mts[k]),
                               selection=None if sel is None else sel[k], invert_selection=invert) for k in order]

    far = sb.scenes.config4_far_to_near(pos)
    assert sorted(far) == list(range(8)) and far[0] in (0, 7) and far[-1] in (3, 4)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    for order in (far, [3, 0, 7, 1, 6, 2, 5, 4]):
        mm.render(target, w, h, order)
        torch.cuda.synchronize()
        oimg, _ = ob.render(omodels(order), ocam, ogt, n_threads=threads)
        d = idiff(target.cpu().numpy(), oimg)
        assert d <= 2, f"order {order}: max-abs {d}/255"
    # per-model artefacts of the last frame
    for k in (0, 5):
        pre = ob.preprocess(omodels([k])[0], ocam, ogt)
        idx, V = mm.read_model_indices(k, pre["count"])
        assert V == pre["count"]
        assert np.array_equal(idx, ob.radix_sort(pre["keys"][:V].view(np.uint32), pre["indices"][:V])[1])
    # rectangle mask, evaluated on the device per model (selection::viewport over the analytic rectangle)
    sel = {}
    for k in range(8):
        mm.select_rect(k, *sb.scenes.CONFIG4_RECT)
        sel[k] = mm.read_selection(k, n)
        assert np.array_equal(sel[k], ob.select_rect(omodels([k])[0], ocam, *sb.scenes.CONFIG4_RECT)), f"model {k}: mask differs"
    assert sum(int(np.unpackbits(s.view(np.uint8)).sum()) for s in sel.values()) > 100_000
    frames = {}
    for invert in (0, 1):
        for k in range(8):
            mm.enable_selection(k, True, bool(invert))
        mm.render(target, w, h, far)
        torch.cuda.synchronize()
        frames[invert] = target.cpu().numpy()
        oimg, _ = ob.render(omodels(far, sel, invert), ocam, ogt, n_threads=threads)
        d = idiff(frames[invert], oimg)
        assert d <= 2, f"invert {invert}: max-abs {d}/255"
    # invert = 0 shows only what projects into the rectangle [480,1440)x[270,810): the band left of it (80 px clear of the
    # border, same rows) is empty, while invert = 1 hides exactly those Gaussians and leaves the band lit
    assert frames[0][300:780, :400, :3].max() == 0 and frames[1][300:780, :400, :3].max() > 0
    assert frames[0][300:780, 600:1300, :3].max() > 0
    mm.close()


# ------------------------------------------------------------------------------------------ config 5b

def test_config5b_8K_strip_against_oracle(sb, ob, ctx):
    """One 7680x4320 frame of the 6 M scene: the strip rank 3 of 8 would rasterize (sharding.strip_rows), rendered through
    SbTarget.row0/rows, against the oracle's render of the same rows; visible set and order as in the full frame."""
    import torch
    from splat_b200 import sharding
    n, w, h = 6_000_000, 7680, 4320
    g = scene(sb, n, sb.scenes.BASE_SEED + 2)
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    om = ob.OracleModel(pods, n)
    ocam, ogt = ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod()
    threads = ob.use_all_host_threads()
    row0, rows = sharding.strip_rows(h, 8, 3)
    assert row0 % 16 == 0 and rows > 0
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    strip = torch.zeros((rows, w, 4), dtype=torch.uint8, device="cuda")
    v.render(strip, w, h, row0=row0, rows=rows)
    torch.cuda.synchronize()
    assert not v.read_frame_stats()["overflowed"]
    check_artifacts(ob, read_artifacts(sb, v, n), ob.preprocess(om, ocam, ogt), n)
    ostrip, _ = ob.render(om, ocam, ogt, row0=row0, rows=rows, n_threads=threads)
    d = idiff(strip.cpu().numpy(), ostrip)
    assert d <= 2, f"8K strip rows [{row0},{row0 + rows}): max-abs {d}/255"
    v.close()
    _SCENES.clear()
