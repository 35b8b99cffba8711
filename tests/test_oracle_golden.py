"""CPU: the oracle against the committed golden vectors and against its numpy twin."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(HERE, "golden", "golden.json")))


def hexes(a):
    return [f"0x{int(x):08x}" for x in np.asarray(a).view(np.uint32)]


def single_scene(ob):
    g = np.zeros(1, dtype=ob.GAUSSIAN_DTYPE)
    g["pos"][0] = (0, 0, 1); g["rot"][0] = (0, 0, 0, 1); g["scale"][0] = (1, 1, 1); g["color"][0] = (255, 0, 0, 255)
    return ob.OracleModel(ob.pack_gaussians(g), 1), ob.camera_pod((0, 0, 0), 0.1, 0.1, 1024, 1024), ob.gaussian_transform_pod()


def ply_scene(ob):
    props = np.load(os.path.join(HERE, "golden", "model_ply_props.npy"))
    gs = ob.gaussians_from_ply_props(props)
    half = np.float32(np.pi) / 2  # examples/simple.rs:168-173: 180 degrees about Z
    mt = ob.model_transform_pod((0, 0, 0), (0, 0, float(np.sin(half)), float(np.cos(half))), (1, 1, 1))
    return gs, ob.OracleModel(ob.pack_gaussians(gs), 9, model_transform=mt), ob.camera_pod((0, 0, 0), 0.0, 0.0, 1280, 720), ob.gaussian_transform_pod()


def ulp_distance(a_hex, b_hex):
    return abs(int(a_hex, 16) - int(b_hex, 16))


def test_survey_kat_single_gaussian(ob, gold):
    m, cam, gt = single_scene(ob)
    kat = gold["survey_kat"]["single"]
    p = ob.preprocess(m, cam, gt)
    assert p["count"] == 1
    # key = 1 - ndc.z: SURVEY's value came from mixed f64/f32 numpy; allow 1 ulp of ndc.z (= 8 key ulps)
    assert ulp_distance(hexes(p["keys"][:1])[0], kat["key"]) <= 8
    s = ob.project(m, cam, gt, p["indices"][:1])[0]
    assert abs(s["cx"] - kat["pixel_centre"][0]) < 0.01 and abs(s["cy"] - kat["pixel_centre"][1]) < 0.01
    assert s["ext_x"] == kat["quad_half_extent"] and s["ext_y"] == kat["quad_half_extent"]  # both axes clamp at 1024
    img, st = ob.render(m, cam, gt)
    assert st["alive_pixels"] == 1024 * 1024  # covers the whole target
    assert img[..., 0].min() > 0 and img[..., 1].max() == 0 and img[..., 2].max() == 0 and img[..., 3].min() == 255


def test_survey_kat_model_ply(ob, gold):
    gs, m, cam, gt = ply_scene(ob)
    kat = gold["survey_kat"]["model_ply"]
    p = ob.preprocess(m, cam, gt)
    vis = p["indices"][: p["count"]].tolist()
    assert vis == kat["visible"] and sorted(set(range(9)) - set(vis)) == kat["culled"]
    keys = dict(zip(vis, hexes(p["keys"][: p["count"]])))
    for k, v in kat["keys"].items():
        assert ulp_distance(keys[int(k)], v) <= 128, (k, keys[int(k)], v)  # <= 2 ulp of ndc.z near 1.0
    # tie structure: a 2-way and a 3-way tie
    assert keys[1] == keys[5] and keys[3] == keys[6] == keys[8] and keys[2] not in (keys[1], keys[3])


def test_oracle_regression_pins(ob, gold):
    o = gold["oracle"]
    m, cam, gt = single_scene(ob)
    assert hexes(np.frombuffer(bytes(cam), dtype=np.uint32)) == o["single"]["camera_pod"]
    p = ob.preprocess(m, cam, gt)
    assert hexes(p["keys"][:1])[0] == o["single"]["key"]
    assert p["draw_args"].tolist() == o["single"]["draw_args"] == [6, 1, 0, 0]
    assert p["sort_args"].tolist() == o["single"]["sort_args"] == [1, 1, 1]
    assert np.all(p["keys"][1:3840] == 2.0)  # post: pad to a whole 3840-key block
    s = ob.project(m, cam, gt, p["indices"][:1])[0]
    for k, v in o["single"]["splat"].items():
        assert float(s[k]) == v, k
    img, _ = ob.render(m, cam, gt)
    assert [int(img[..., c].astype(np.int64).sum()) for c in range(4)] == o["single"]["image_sum_rgba"]
    assert img[600, 600].tolist() == o["single"]["pixel_600_600"] and img[0, 0].tolist() == o["single"]["pixel_0_0"]

    gs, m, cam, gt = ply_scene(ob)
    assert gs["color"].tolist() == o["model_ply"]["colors"]
    p = ob.preprocess(m, cam, gt)
    V = p["count"]
    assert V == o["model_ply"]["count"] and hexes(p["mask"]) == o["model_ply"]["mask"]
    assert p["indices"][:V].tolist() == o["model_ply"]["indices_compacted"]
    assert hexes(p["keys"][:V]) == o["model_ply"]["keys_compacted"]
    sk, si = ob.radix_sort(p["keys"][:V].view(np.uint32), p["indices"][:V])
    assert si.tolist() == o["model_ply"]["indices_sorted"] and hexes(sk) == o["model_ply"]["keys_sorted"]
    img, st = ob.render(m, cam, gt)
    assert [int(img[..., c].astype(np.int64).sum()) for c in range(4)] == o["model_ply"]["image_sum_rgba"]
    assert st["alive_pixels"] == o["model_ply"]["alive_pixels"]
    for k, v in o["strides"].items():
        sh, cov = map(int, k.split(","))
        assert ob.pod_stride(sh, cov) == v
    for x, v in o["exp_neg_poly"].items():
        assert hexes(np.array([ob.exp_neg_poly(float(x))], dtype=np.float32))[0] == v


def test_exp_neg_poly_accuracy(ob):
    xs = np.linspace(0, 12, 2001)
    got = np.array([ob.exp_neg_poly(float(x)) for x in xs])
    ref = np.exp(-xs.astype(np.float32).astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 2e-6


def test_radix_sort_is_stable_lsd(ob):
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 255, 256, 257, 3840, 50_000):
        k = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        k[: n // 2] &= np.uint32(0xFFFF0000)
        v = np.arange(n, dtype=np.uint32)
        sk, sv = ob.radix_sort(k, v)
        order = np.argsort(k, kind="stable")
        assert np.array_equal(sk, k[order]) and np.array_equal(sv, v[order])


@pytest.mark.parametrize("camera", ["outside", "inside"])
def test_c_oracle_matches_numpy_twin(ob, sb, camera):
    """Two independent statements of the arithmetic contract agree bit for bit (mask + keys)."""
    from oracle import oracle_np
    n = 40_000
    g = sb.scenes.synthetic_gaussians(n, 17)
    pods = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE))
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE if camera == "outside" else sb.scenes.CAMERA_INSIDE
    cam = ob.camera_pod(pos, yaw, pitch, 1280, 720)
    q = np.array([0.2, -0.3, 0.1, 0.9], dtype=np.float32)
    q /= np.linalg.norm(q)
    mt = ob.model_transform_pod((0.5, -1.0, 2.0), tuple(q), (1.1, 0.9, 1.3))
    gt = ob.gaussian_transform_pod(1.0, 0, 3, False, 3.0)
    rng = np.random.default_rng(2)
    sel = rng.integers(0, 2**32, size=(n + 31) // 32, dtype=np.uint64).astype(np.uint32)
    for selection, invert in ((None, 1), (sel, 1), (sel, 0)):
        m = ob.OracleModel(pods, n, model_transform=mt, selection=selection, invert_selection=invert)
        p = ob.preprocess(m, cam, gt)
        cov3d = np.frombuffer(pods, dtype=np.float32).reshape(n, 56)[:, 49:55]
        vis, keys, _ = oracle_np.preprocess(g["pos"], cov3d, cam, mt, gt, selection, invert)
        idx = np.nonzero(vis)[0].astype(np.uint32)
        assert p["count"] == len(idx)
        assert np.array_equal(p["indices"][: p["count"]], idx)
        assert np.array_equal(p["keys"][: p["count"]].view(np.uint32), keys[idx].view(np.uint32))
        if camera == "inside":
            assert 0 < len(idx) < n // 2


def test_pad_and_args_semantics(ob, sb):
    """post (preprocess.wesl:108-126): x = ceil(V/3840); keys [V, x*3840) = 2.0."""
    n = 10_000
    g = sb.scenes.synthetic_gaussians(n, 3)
    m = ob.OracleModel(ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE)), n)
    cam = ob.camera_pod(*sb.scenes.CAMERA_OUTSIDE, 640, 360)
    p = ob.preprocess(m, cam, ob.gaussian_transform_pod())
    V = p["count"]
    assert p["draw_args"].tolist() == [6, V, 0, 0] and p["sort_args"].tolist() == [(V + 3839) // 3840, 1, 1]
    assert np.all(p["keys"][V:(V + 3839) // 3840 * 3840] == 2.0)
    assert np.all(p["keys"][:V] >= 0) and np.all(p["keys"][:V] <= 1)


def test_selection_semantics(ob):
    """invert = 1 (default): set bits are hidden; invert = 0: only set bits are shown."""
    g = np.zeros(64, dtype=ob.GAUSSIAN_DTYPE)
    g["pos"][:, 2] = 5.0
    g["pos"][:, 0] = np.linspace(-1, 1, 64)
    g["rot"][:, 3] = 1.0
    g["scale"][:] = 0.01
    g["color"][:] = 255
    pods = ob.pack_gaussians(g)
    cam = ob.camera_pod((0, 0, 0), 0.0, 0.0, 256, 256)
    gt = ob.gaussian_transform_pod()
    sel = np.array([0x0000FFFF, 0x80000001], dtype=np.uint32)
    assert ob.preprocess(ob.OracleModel(pods, 64, selection=np.zeros(2, np.uint32), invert_selection=1), cam, gt)["count"] == 64
    hid = ob.preprocess(ob.OracleModel(pods, 64, selection=sel, invert_selection=1), cam, gt)
    assert hid["indices"][: hid["count"]].tolist() == [i for i in range(64) if not (sel[i // 32] >> (i % 32)) & 1]
    only = ob.preprocess(ob.OracleModel(pods, 64, selection=sel, invert_selection=0), cam, gt)
    assert only["indices"][: only["count"]].tolist() == [i for i in range(64) if (sel[i // 32] >> (i % 32)) & 1]


def test_multi_model_order_matters(ob, sb):
    """multi_model.rs:505-527: model k+1 is composited over model k (no depth merge)."""
    def blob(color, z):
        g = np.zeros(1, dtype=ob.GAUSSIAN_DTYPE)
        g["pos"][0] = (0, 0, z); g["rot"][0] = (0, 0, 0, 1); g["scale"][0] = (0.3, 0.3, 0.3); g["color"][0] = color
        return ob.OracleModel(ob.pack_gaussians(g), 1)
    red, green = blob((255, 0, 0, 255), 3.0), blob((0, 255, 0, 255), 6.0)
    cam = ob.camera_pod((0, 0, 0), 0.0, 0.0, 128, 128)
    gt = ob.gaussian_transform_pod()
    a, _ = ob.render([red, green], cam, gt)
    b, _ = ob.render([green, red], cam, gt)
    c = 64
    assert a[c, c, 1] > a[c, c, 0]      # green drawn last wins although it is farther away
    assert b[c, c, 0] > b[c, c, 1]


def test_unorm8_newton_identity(ob):
    """The kernel's division-free u8/255 is exact for every byte (see sb_preprocess.cu: unorm8)."""
    assert ob.lib().so_unorm8_newton_mismatches() == 0


def test_oracle_brush_mask_is_a_capsule():
    """so_select_brush: a Gaussian is selected iff the texel under its centre lies within `radius` of the stroke."""
    from oracle import binding as ob
    import numpy as np
    g = np.zeros(4, dtype=ob.GAUSSIAN_DTYPE)
    g["rot"][:, 3] = 1
    g["scale"][:] = 0.01
    g["color"][:] = 255
    # camera at the origin looking down +z (yaw = pitch = 0): x right, y up; 200x100 target
    g["pos"] = [(0, 0, 5), (0.5, 0, 5), (0, 0.5, 5), (0, 0, -5)]
    pods = ob.pack_gaussians(g, 0, 0)
    m = ob.OracleModel(pods, 4)
    cam = ob.camera_pod((0.0, 0.0, 0.0), 0.0, 0.0, 200, 100)
    centre = ob.select_brush(m, cam, [(100.0, 50.0)], 3.0)
    assert centre[0] == 0b0001  # only the Gaussian under the dab; the one behind the camera is culled
    stroke = ob.select_brush(m, cam, [(90.0, 50.0), (140.0, 50.0)], 1.0)
    assert stroke[0] & 0b0001 and not stroke[0] & 0b0100 and not stroke[0] & 0b1000
    both = ob.select_brush(m, cam, [(100.0, 0.0), (100.0, 99.0)], 1.0, accumulate=True, dest=stroke)
    assert both[0] & 0b0100 and both[0] & stroke[0] == stroke[0]


def test_oracle_draw_equals_render_and_clips_whole_quads(ob, sb):
    """so_draw (Renderer<G, ()>::render on caller indices) fed with the preprocess + sort output reproduces so_render, and
    drops the quads whose centre has w <= 0 or ndc z outside [0, 1] when it is handed indices no preprocess produced."""
    n, w, h = 4000, 320, 180
    g = sb.scenes.synthetic_gaussians(n, 5)
    pods = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE))
    pos, yaw, pitch = sb.scenes.CAMERA_INSIDE
    cam, gt = ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod()
    om = ob.OracleModel(pods, n)
    pre = ob.preprocess(om, cam, gt)
    V = pre["count"]
    _, idx = ob.radix_sort(pre["keys"][:V].view(np.uint32), pre["indices"][:V])
    ref, _ = ob.render(om, cam, gt, strict_exp=True)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    ob.draw(om, cam, gt, idx, out, strict_exp=True)
    assert np.array_equal(out, ref)
    # every Gaussian in index order: the culled ones behind the camera must not draw, the ones beside the frustum may
    sp = ob.project(om, cam, gt, np.arange(n, dtype=np.uint32))
    behind = ~((sp["z"] >= 0) & (sp["z"] <= 1))
    assert behind.sum() > n // 4
    only_behind = np.flatnonzero(behind).astype(np.uint32)
    out2 = np.full((h, w, 4), 7, dtype=np.uint8)
    ob.draw(om, cam, gt, only_behind, out2, strict_exp=True)
    assert (out2[..., :3] == 0).all() and (out2[..., 3] == 255).all()  # cleared to BLACK, nothing drawn


def test_alpha_cutoff_identity_bound():
    """The rounding argument behind the exact alpha cut-off (sb_common.cuh kAlphaCut = 0.00195): for every destination
    d in 0..255, source colours c255 in [0, 255] and alpha <= kAlphaCut, rint(fma(d, 1 - alpha, c255 * alpha)) == d in the
    f32 arithmetic the blend uses — and the no-discard build's premise exp(-6.3) < kAlphaCut."""
    f32 = np.float32
    cut = f32(0.00195)
    assert np.exp(-6.3) < float(cut) * 0.99
    rng = np.random.default_rng(1)
    d = np.arange(256, dtype=np.float64)[:, None, None]
    c = np.concatenate([np.array([0.0, 255.0, 254.99998, 1e-3]), rng.uniform(0, 255, 60)]).astype(f32).astype(np.float64)[None, :, None]
    alphas = np.concatenate([np.array([cut, np.nextafter(cut, f32(0)), f32(1e-30), f32(0)], dtype=f32),
                             (cut * rng.uniform(0, 1, 200).astype(f32)).astype(f32)])
    om = (f32(1.0) - alphas).astype(f32).astype(np.float64)[None, None, :]          # 1 - alpha, rounded to f32
    ca = (c.astype(f32) * alphas[None, None, :]).astype(f32).astype(np.float64)    # c255 * alpha, rounded to f32
    x = (d * om + ca).astype(f32)                                                   # the fma: one rounding (f64 sum is exact enough: 1e-13)
    assert np.array_equal(np.rint(x.astype(np.float64)), np.broadcast_to(d, x.shape))
    # and just above the bound the identity does fail somewhere: the cut is not vacuous by orders of magnitude
    a_big = f32(0.0021)
    x2 = (d[:, 0, 0] * float(f32(1.0) - a_big) + float(f32(255.0) * a_big)).astype(f32)
    assert (np.rint(x2) != d[:, 0, 0]).any()


def test_srgb_tables_and_oracle_srgb_target(ob, sb):
    """The shared sRGB tables: every code decodes inside its own encode interval (a blend with alpha = 0 is the identity), both
    generated copies are identical and up to date, thresholds increase; and the oracle's Rgba8UnormSrgb frame agrees with a
    float64 decode-blend-encode restatement of the same composite."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    a = open(os.path.join(root, "oracle", "srgb_tables.h")).read()
    b = open(os.path.join(root, "wgpu-3dgs-viewer_b200", "csrc", "sb_srgb_tables.h")).read()
    bits = lambda txt: [int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)]
    assert bits(a) == bits(b) and len(bits(a)) == 512
    dec = np.array(bits(a)[:256], dtype=np.uint32).view(np.float32)
    thr = np.array(bits(a)[256:], dtype=np.uint32).view(np.float32)
    assert np.all(np.diff(thr.astype(np.float64)) > 0) and dec[0] == 0.0 and dec[255] == 1.0
    assert np.all(thr <= dec) and np.all(dec[:-1] < thr[1:])
    x = np.arange(256) / 255.0
    ref = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    assert np.allclose(dec, ref, rtol=1e-6, atol=1e-9)

    n, w, h = 4000, 320, 180
    g = sb.scenes.synthetic_gaussians(n, 5)
    pods = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE))
    m = ob.OracleModel(pods, n)
    cam = ob.camera_pod(*sb.scenes.CAMERA_OUTSIDE, w, h)
    gt = ob.gaussian_transform_pod()
    img, _ = ob.render(m, cam, gt, ob.TARGET_RGBA8_SRGB)
    imgb, _ = ob.render(m, cam, gt, ob.TARGET_BGRA8_SRGB)
    assert np.array_equal(img[..., [2, 1, 0, 3]], imgb)
    # float64 restatement: the oracle's splat records composited with decode -> blend -> encode per blend
    p = ob.preprocess(m, cam, gt)
    _, idx = ob.radix_sort(p["keys"][: p["count"]].view(np.uint32), p["indices"][: p["count"]])
    sp = ob.project(m, cam, gt, idx)
    acc = np.zeros((h, w, 3))
    eotf = lambda c: np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    oetf = lambda l: np.where(l <= 0.0031308, 12.92 * l, 1.055 * np.power(np.maximum(l, 0), 1 / 2.4) - 0.055)
    for s_ in sp:
        if not s_["valid"]:
            continue
        x0, x1 = int(max(np.floor(s_["cx"] - s_["ext_x"] - 1), 0)), int(min(np.ceil(s_["cx"] + s_["ext_x"] + 1), w - 1))
        y0, y1 = int(max(np.floor(s_["cy"] - s_["ext_y"] - 1), 0)), int(min(np.ceil(s_["cy"] + s_["ext_y"] + 1), h - 1))
        if x1 < x0 or y1 < y0:
            continue
        dx = (np.arange(x0, x1 + 1) + 0.5) - float(s_["cx"])
        dy = ((np.arange(y0, y1 + 1) + 0.5) - float(s_["cy"]))[:, None]
        qx, qy = dx * float(s_["ax"]) + dy * float(s_["ay"]), dx * float(s_["bx"]) + dy * float(s_["by"])
        r2 = qx * qx + qy * qy
        alpha = np.where(r2 <= 9.0, float(s_["a"]) * np.exp(-r2), 0.0)[..., None]
        src = np.minimum(np.array([s_["r"], s_["g"], s_["b"]], dtype=np.float64), 1.0)
        win = acc[y0:y1 + 1, x0:x1 + 1]
        win[...] = np.rint(np.clip(oetf(eotf(win / 255.0) * (1 - alpha) + src * alpha), 0, 1) * 255.0)
    d = np.abs(acc - img[..., :3].astype(np.float64))
    assert d.max() <= 2 and np.count_nonzero(d) < 0.01 * d.size
