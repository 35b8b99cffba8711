"""CPU: the C oracle's DRAW stage against its two independent numpy statements (oracle/oracle_np.py, second half):
  * project_f32 / composite_f32 — strict-f32 contract, bit for bit;
  * draw_reference_f64 — written from render.wesl / utils.wesl alone in float64 and in the shader's clip-space-quad
    formulation: within the framebuffer tolerance, and exactly equal on almost every pixel.
Scenes: BASELINE config 1 (coverage/model.ply under the examples/simple defaults) and a 20 K synthetic scene."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("cx", "cy", "ax", "ay", "bx", "by", "r", "g", "b", "a", "ext_x", "ext_y", "z")


def _ply(ob):
    props = np.load(os.path.join(HERE, "golden", "model_ply_props.npy"))
    gs = ob.gaussians_from_ply_props(props)
    half = np.float32(np.pi) / 2
    mt = ob.model_transform_pod((0, 0, 0), (0, 0, float(np.sin(half)), float(np.cos(half))), (1, 1, 1))
    return ob.pack_gaussians(gs), 9, mt, ob.camera_pod((0, 0, 0), 0.0, 0.0, 1280, 720)


def _sorted_visible(ob, model, cam, gt):
    p = ob.preprocess(model, cam, gt)
    V = p["count"]
    _, idx = ob.radix_sort(p["keys"][:V].view(np.uint32), p["indices"][:V])
    return idx


def _assert_splats_equal(c, t):
    valid = c["valid"].astype(bool)
    assert np.array_equal(valid, np.asarray(t["valid"], dtype=bool))
    for k in FIELDS:
        a, b = c[k][valid].view(np.uint32), np.asarray(t[k], dtype=np.float32)[valid].view(np.uint32)
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, f"{k}: {len(bad)} of {valid.sum()} differ, first at {bad[:3]}"


def test_fma32_is_exact():
    """The fma emulation against exact rational arithmetic on adversarial operands (sums landing near rounding boundaries)."""
    from fractions import Fraction
    from oracle import oracle_np
    rng = np.random.default_rng(11)
    a = rng.standard_normal(4000).astype(np.float32)
    b = rng.standard_normal(4000).astype(np.float32)
    c = (-(a.astype(np.float64) * b.astype(np.float64))).astype(np.float32)  # cancellation: exposes double rounding
    c[::2] = rng.standard_normal(2000).astype(np.float32)
    # force ties: c = 2^24 + 1 (odd integer above f32's precision needs the product's sticky bits)
    a[:200] = np.float32(1.0) + np.float32(2.0 ** -23) * rng.integers(0, 64, 200).astype(np.float32)
    b[:200] = np.float32(2.0 ** -25) * (1 + rng.integers(0, 3, 200)).astype(np.float32)
    c[:200] = np.float32(1.0) + np.float32(2.0 ** -23) * rng.integers(0, 64, 200).astype(np.float32)
    got = oracle_np.fma32(a, b, c)
    for i in range(len(a)):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        # correctly rounded f32 of an exact rational: compare against both neighbours of the candidate
        r = np.float32(float(exact))  # float(Fraction) is correctly rounded to f64; may double-round: fix by neighbours
        cands = [np.nextafter(r, np.float32(-np.inf)), r, np.nextafter(r, np.float32(np.inf))]
        best = min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(np.float32(v).view(np.uint32)) & 1))
        assert got[i].view(np.uint32) == np.float32(best).view(np.uint32), (i, a[i], b[i], c[i])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_vertex_stage_twin_model_ply(ob, mode):
    from oracle import oracle_np
    pods, n, mt, cam = _ply(ob)
    gt = ob.gaussian_transform_pod(1.0, mode, 3, False, 3.0)
    m = ob.OracleModel(pods, n, model_transform=mt)
    idx = _sorted_visible(ob, m, cam, gt)
    _assert_splats_equal(ob.project(m, cam, gt, idx), oracle_np.project_f32(pods, n, 0, 0, cam, mt, gt, idx))


@pytest.mark.parametrize("sh_fmt,cov_fmt,sh_deg,no_sh0", [(0, 0, 3, False), (1, 1, 3, False), (2, 1, 2, False), (0, 0, 1, True), (3, 0, 3, False)])
def test_vertex_stage_twin_20k(ob, sb, sh_fmt, cov_fmt, sh_deg, no_sh0):
    """view_color (all SH degrees / pod formats), cov2d_axes -> inverse axes, extents, centre: bit for bit on 20 K Gaussians
    under a non-trivial model transform."""
    from oracle import oracle_np
    n = 20_000
    g = sb.scenes.synthetic_gaussians(n, 23)
    pods = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE), sh_fmt, cov_fmt)
    q = np.array([0.1, 0.25, -0.2, 0.94], dtype=np.float32)
    q /= np.linalg.norm(q)
    mt = ob.model_transform_pod((0.3, -0.7, 1.5), tuple(q), (1.2, 0.8, 1.1))
    cam = ob.camera_pod(*sb.scenes.CAMERA_OUTSIDE, 960, 540)
    gt = ob.gaussian_transform_pod(1.3, 0, sh_deg, no_sh0, 2.5)
    m = ob.OracleModel(pods, n, sh_fmt, cov_fmt, model_transform=mt)
    idx = _sorted_visible(ob, m, cam, gt)
    assert len(idx) > n // 2
    _assert_splats_equal(ob.project(m, cam, gt, idx), oracle_np.project_f32(pods, n, sh_fmt, cov_fmt, cam, mt, gt, idx))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_composite_twin_bit_exact_20k(ob, sb, mode):
    """frag_main + per-blend re-quantised ALPHA_BLENDING, strict contract: the numpy composite of the C oracle's own splat
    records equals the C oracle's frame bit for bit (strict exp)."""
    from oracle import oracle_np
    n, w, h = 20_000, 480, 270
    g = sb.scenes.synthetic_gaussians(n, 29)
    pods = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE))
    cam = ob.camera_pod(*sb.scenes.CAMERA_OUTSIDE, w, h)
    gt = ob.gaussian_transform_pod(1.0, mode, 3, False, 3.0)
    m = ob.OracleModel(pods, n)
    idx = _sorted_visible(ob, m, cam, gt)
    sp = ob.project(m, cam, gt, idx)
    img, _ = ob.render(m, cam, gt, strict_exp=True)
    twin = oracle_np.composite_f32({k: sp[k] for k in sp.dtype.names}, w, h, mode, 3.0, strict_exp=True)
    assert np.array_equal(img, twin), f"{np.count_nonzero(img != twin)} channel values differ"


@pytest.mark.parametrize("scene,mode", [("ply", 0), ("ply", 1), ("20k", 0), ("20k", 1), ("20k", 2)])
def test_draw_stage_against_f64_wgsl_reference(ob, sb, scene, mode):
    """The whole draw stage against the float64 restatement written from the WGSL text (clip-space quad, interpolated
    quad_offset, true exp): max-abs <= 2/255 in splat mode, and the two agree exactly on >= 99 % of the channel values."""
    from oracle import oracle_np
    if scene == "ply":
        pods, n, mt, cam = _ply(ob)
        w, h = 1280, 720
    else:
        n, w, h = 20_000, 480, 270
        g = sb.scenes.synthetic_gaussians(n, 31)
        pods = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE))
        mt = ob.model_transform_pod((0.2, 0.1, -0.4), (0.0, 0.0871557, 0.0, 0.9961947), (1.0, 1.1, 0.9))
        cam = ob.camera_pod(*sb.scenes.CAMERA_OUTSIDE, w, h)
    gt = ob.gaussian_transform_pod(1.0, mode, 3, False, 3.0)
    m = ob.OracleModel(pods, n, model_transform=mt)
    idx = _sorted_visible(ob, m, cam, gt)
    img, _ = ob.render(m, cam, gt, strict_exp=False)
    ref = oracle_np.draw_reference_f64(pods, n, 0, 0, cam, mt, gt, idx, w, h)
    d = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    if mode == 0:
        assert d.max() <= 2, f"max-abs {d.max()}/255"
    else:
        # ellipse / point fragments are step functions of r^2 (outline ring, discard edge, quad edge): a pixel centre that sits
        # on a step within f32 rounding lands on either side, which no tolerance on the colour can absorb — bound how many do
        assert np.count_nonzero(d.max(axis=-1) > 2) <= 5e-4 * d.shape[0] * d.shape[1], f"{np.count_nonzero(d.max(axis=-1) > 2)} pixels off by > 2/255"
    assert np.count_nonzero(d) <= 0.01 * d.size, f"{np.count_nonzero(d)} of {d.size} channel values differ"
    assert img[..., :3].max() > 0
