"""The reference's own e2e tests (tests/e2e/viewer.rs, multi_model.rs, selection.rs) restated
through the C ABI: presence/absence assertions on the rendered target."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W = H = 1024  # tests/common/given.rs:22-37


def given_camera(sb):
    return sb.camera_pod((0, 0, 0), 0.1, 0.1, W, H)  # tests/common/given.rs:6-20


def render(sb, ctx, g, *, transform=None, gt=None, pod_path=False):
    import torch
    v = sb.Viewer(ctx, gaussians=g)
    cam = given_camera(sb)
    if pod_path:
        v.update_camera_with_pod(cam)
    else:
        v.update_camera((0, 0, 0), 0.1, 0.1, W, H)
    if transform:
        v.update_model_transform(*transform)
    if gt:
        v.update_gaussian_transform(*gt)
    t = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    v.render(t, W, H)
    torch.cuda.synchronize()
    out = t.cpu().numpy()
    v.close()
    return out


def sums(img):
    return [(img[..., c] > 0).sum() for c in range(4)]


def test_viewer_render_red_gaussian(sb, ctx):  # tests/e2e/viewer.rs:70-95
    r, g, b, a = sums(render(sb, ctx, sb.scenes.single_red_gaussian()))
    assert r > 1 and g < 1 and b < 1 and a > 1


def test_viewer_pod_and_non_pod_paths_equal(sb, ctx):  # tests/e2e/viewer.rs:39-68
    a = render(sb, ctx, sb.scenes.single_red_gaussian(), pod_path=True)
    b = render(sb, ctx, sb.scenes.single_red_gaussian(), pod_path=False)
    assert np.array_equal(a, b)


def test_viewer_no_sh0_renders_grey(sb, ctx):  # tests/e2e/viewer.rs:97-124
    r, g, b, a = sums(render(sb, ctx, sb.scenes.single_red_gaussian(), gt=(1.0, 0, 3, True, 3.0)))
    assert r > 1 and g > 1 and b > 1 and a > 1


def test_viewer_model_behind_camera_draws_nothing(sb, ctx):  # tests/e2e/viewer.rs:149-182
    img = render(sb, ctx, sb.scenes.single_red_gaussian(), transform=((0, 0, -2), (0, 0, 0, 1), (1, 1, 1)))
    r, g, b, a = sums(img)
    assert r == 0 and g == 0 and b == 0


def test_gaussians_buffer_size(sb, ctx):  # tests/e2e/multi_model.rs:43-53
    g = sb.scenes.synthetic_gaussians(100, 1)
    v = sb.Viewer(ctx, gaussians=g)
    assert v.device_pointers()["gaussians"][1] == 100 * sb.pod_stride(0, 0) == 100 * 224
    ptrs = v.device_pointers()
    assert ptrs["indirect_args"][1] == 16 and ptrs["radix_sort_indirect_args"][1] == 12  # tests/buffer/indirect_args.rs
    assert ptrs["indirect_indices"][1] == 100 * 4
    v.close()


def test_multi_model_red_and_green(sb, ctx):  # tests/e2e/multi_model.rs:100-154, 330-373
    import torch
    red = sb.scenes.single_red_gaussian()
    green = red.copy()
    green["color"][0] = (0, 255, 0, 255)
    green["pos"][0] = (0.5, 0, 1)
    green["scale"][0] = (0.05, 0.05, 0.05)
    red["scale"][0] = (0.05, 0.05, 0.05)
    mm = sb.MultiModelViewer(ctx)
    mm.update_camera_with_pod(given_camera(sb))
    mm.insert_model(1, sb.pack_gaussians(red), 1)
    mm.insert_model(2, sb.pack_gaussians(green), 1)
    t = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    mm.render(t, W, H, [1, 2])
    torch.cuda.synchronize()
    r, g, b, a = sums(t.cpu().numpy())
    assert r > 1 and g > 1 and b < 1
    mm.remove_model(2)
    mm.render(t, W, H, [1])
    torch.cuda.synchronize()
    r, g, b, a = sums(t.cpu().numpy())
    assert r > 1 and g < 1
    mm.close()


def _mm_reference_scene(sb, ctx):
    """tests/e2e/multi_model.rs: "red" = one unit Gaussian at (0,0,1), "green" = one at (1,0,1), given::camera_pod()."""
    red = sb.scenes.single_red_gaussian()
    green = red.copy()
    green["color"][0] = (0, 255, 0, 255)
    green["pos"][0] = (1, 0, 1)
    mm = sb.MultiModelViewer(ctx)
    mm.insert_model_from_gaussians(1, red)      # viewer.insert_model(&device, "red", &red_gaussians)
    mm.insert_model_from_gaussians(2, green)
    return mm


def _mm_pixel_sums(mm, keys):
    import torch
    t = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    mm.render(t, W, H, keys)
    torch.cuda.synchronize()
    img = t.cpu().numpy()
    return img, [int(img[..., c].astype(np.int64).sum()) for c in range(4)]


def test_multi_model_update_camera_with_or_without_pod_equal(sb, ctx):  # tests/e2e/multi_model.rs:56-112
    a = _mm_reference_scene(sb, ctx)
    a.update_camera((0, 0, 0), 0.1, 0.1, W, H)                  # update_camera(&queue, &given::camera(), size)
    b = _mm_reference_scene(sb, ctx)
    b.update_camera_with_pod(given_camera(sb))                  # update_camera_with_pod(&queue, &given::camera_pod())
    img1, _ = _mm_pixel_sums(a, [1, 2])
    img2, _ = _mm_pixel_sums(b, [1, 2])
    assert np.array_equal(img1, img2)
    a.close()
    b.close()


def test_multi_model_render_should_render_correctly(sb, ctx):  # tests/e2e/multi_model.rs:114-155
    mm = _mm_reference_scene(sb, ctx)
    mm.update_camera_with_pod(given_camera(sb))
    _, (r, g, b, a) = _mm_pixel_sums(mm, [1, 2])
    assert r > 1 and g > 1 and b < 1 and a > 1
    mm.close()


@pytest.mark.parametrize("with_pod", [False, True])
def test_multi_model_no_sh0_renders_grey(sb, ctx, with_pod):  # tests/e2e/multi_model.rs:157-232
    mm = _mm_reference_scene(sb, ctx)
    mm.update_camera_with_pod(given_camera(sb))
    if with_pod:
        mm.update_gaussian_transform_with_pod(sb.gaussian_transform_pod(1.0, sb.MODE_SPLAT, 3, True, 3.0))
    else:
        mm.update_gaussian_transform(1.0, sb.MODE_SPLAT, 3, True, 3.0)
    _, (r, g, b, a) = _mm_pixel_sums(mm, [1, 2])
    assert r > 1 and g > 1 and b > 1 and a > 1
    mm.close()


@pytest.mark.parametrize("with_pod", [False, True])
def test_multi_model_behind_camera_draws_nothing(sb, ctx, with_pod):  # tests/e2e/multi_model.rs:234-328
    mm = _mm_reference_scene(sb, ctx)
    mm.update_camera_with_pod(given_camera(sb))
    for key in (1, 2):
        if with_pod:
            mm.update_model_transform_with_pod(key, sb.model_transform_pod((0, 0, -1), (0, 0, 0, 1), (1, 1, 1)))
        else:
            mm.update_model_transform(key, (0, 0, -1), (0, 0, 0, 1), (1, 1, 1))
    _, (r, g, b, a) = _mm_pixel_sums(mm, [1, 2])
    assert r == 0 and g == 0 and b == 0
    with pytest.raises(sb.SplatError):                           # MultiModelViewerAccessError::ModelNotFound
        mm.update_model_transform(3, (0, 0, 0), (0, 0, 0, 1), (1, 1, 1))
    mm.close()


def test_multi_model_remove_model_should_not_render_removed_model(sb, ctx):  # tests/e2e/multi_model.rs:330-373
    mm = _mm_reference_scene(sb, ctx)
    mm.update_camera_with_pod(given_camera(sb))
    assert mm.remove_model(2)
    _, (r, g, b, a) = _mm_pixel_sums(mm, [1])
    assert r > 1 and g < 1 and b < 1 and a > 1
    with pytest.raises(sb.SplatError):
        mm.render(__import__("torch").zeros((H, W, 4), dtype=__import__("torch").uint8, device="cuda"), W, H, [1, 2])
    mm.close()


def test_selection_default_invert_hides_nothing(sb, ctx):  # src/selection/buffer.rs:157-165
    import torch
    v = sb.Viewer(ctx, gaussians=sb.scenes.single_red_gaussian())
    v.update_camera_with_pod(given_camera(sb))
    v.enable_selection(True)  # empty mask, invert = 1 (default)
    t = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    v.render(t, W, H)
    torch.cuda.synchronize()
    assert sums(t.cpu().numpy())[0] > 1
    v.set_selection(np.array([1], dtype=np.uint32))  # selected + inverted => hidden
    v.render(t, W, H)
    torch.cuda.synchronize()
    assert sums(t.cpu().numpy())[0] == 0
    v.set_invert_selection(False)  # show only selected
    v.render(t, W, H)
    torch.cuda.synchronize()
    assert sums(t.cpu().numpy())[0] > 1
    v.close()


def select_modify_render(sb, ctx, select):
    """tests/e2e/selection.rs:16-116: viewport selection -> NonDestructiveModifier (rgb override to blue) -> render."""
    import torch
    v = sb.Viewer(ctx, gaussians=sb.scenes.single_red_gaussian())
    v.update_camera_with_pod(given_camera(sb))
    select(v)
    v.apply_rgb_override((0.0, 0.0, 1.0), alpha=1.0)
    t = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    v.render(t, W, H)
    torch.cuda.synchronize()
    out = t.cpu().numpy()
    # non-destructive: restoring brings the red Gaussian back
    v.restore_gaussians()
    v.render(t, W, H)
    torch.cuda.synchronize()
    r, g, b, a = sums(t.cpu().numpy())
    assert r > 1 and g < 1 and b < 1 and a > 1
    v.close()
    return sums(out)


BRUSH_RADIUS = 50.0  # ViewportSelector::DEFAULT_BRUSH_RADIUS, src/selection/viewport_selector.rs:201


def test_selected_rectangle_is_modified(sb, ctx):  # tests/e2e/selection.rs:118-137
    r, g, b, a = select_modify_render(sb, ctx, lambda v: v.select_rect(256.0, 256.0, 1024.0 - 256.0, 1024.0 - 256.0))
    assert r < 1 and g < 1 and b > 1 and a > 1


def test_unselected_rectangle_is_not_modified(sb, ctx):  # tests/e2e/selection.rs:139-158
    r, g, b, a = select_modify_render(sb, ctx, lambda v: v.select_rect(0.0, 0.0, 256.0, 256.0))
    assert r > 1 and g < 1 and b < 1 and a > 1


def test_selected_brush_is_modified(sb, ctx):  # tests/e2e/selection.rs:160-179
    r, g, b, a = select_modify_render(sb, ctx, lambda v: v.select_brush([(256.0, 256.0), (768.0, 768.0)], BRUSH_RADIUS))
    assert r < 1 and g < 1 and b > 1 and a > 1


def test_unselected_brush_is_not_modified(sb, ctx):  # tests/e2e/selection.rs:181-196
    r, g, b, a = select_modify_render(sb, ctx, lambda v: v.select_brush([(0.0, 0.0), (256.0, 256.0)], BRUSH_RADIUS))
    assert r > 1 and g < 1 and b < 1 and a > 1


def test_zero_radius_brush_selects_nothing(sb, ctx):  # tests/e2e/selection.rs:198-214
    r, g, b, a = select_modify_render(sb, ctx, lambda v: v.select_brush([(256.0, 256.0), (768.0, 768.0)], 0.0))
    assert r > 1 and g < 1 and b < 1 and a > 1


def test_cleared_selection_is_not_modified(sb, ctx):  # tests/e2e/selection.rs:216-236
    def sel(v):
        v.select_rect(256.0, 256.0, 768.0, 768.0)
        v.set_selection(np.zeros(1, dtype=np.uint32))  # selector.clear
    r, g, b, a = select_modify_render(sb, ctx, sel)
    assert r > 1 and g < 1 and b < 1 and a > 1


def test_rgb_override_matches_oracle_on_edited_pods(sb, ob, ctx):
    """The edit is exactly a rewrite of the colour words: frames equal the oracle's on pods edited the same way on the host."""
    import torch
    n, w, h = 20000, 640, 360
    g = sb.scenes.synthetic_gaussians(n, 31)
    pods = sb.pack_gaussians(g, 1, 1)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n, sh_fmt=1, cov_fmt=1)
    v.update_camera(pos, yaw, pitch, w, h)
    v.set_strict_exp(True)
    v.select_rect(200.0, 100.0, 450.0, 300.0)
    sel = v.read_selection()
    v.apply_rgb_override((0.25, 1.0, 0.5), alpha=0.5)
    t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(t, w, h)
    torch.cuda.synchronize()
    stride = sb.pod_stride(1, 1)
    edited = pods.copy().reshape(n, stride)
    bits = ((sel[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
    assert 0 < bits.sum() < n
    a = np.float32(edited[bits, 15].astype(np.float32) / np.float32(255.0)) * np.float32(0.5)
    edited[bits, 12] = np.uint8(np.rint(np.float32(0.25) * np.float32(255)))
    edited[bits, 13] = 255
    edited[bits, 14] = np.uint8(np.rint(np.float32(0.5) * np.float32(255)))
    edited[bits, 15] = np.rint(a * np.float32(255)).astype(np.uint8)
    oimg, _ = ob.render(ob.OracleModel(edited.reshape(-1), n, 1, 1), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
    assert np.array_equal(t.cpu().numpy(), oimg)
    v.close()


def _basic_color_np(c_u8, rgb=None, hsv=(0.0, 1.0, 1.0), alpha=1.0, contrast=0.0, exposure=0.0, gamma=1.0):
    """numpy restatement of the BasicColorModifiers contract (include/splat_b200.h): every step an individually rounded f32 op."""
    f = np.float32
    c = c_u8.astype(f) / f(255.0)
    r, g, b = c[:, 0], c[:, 1], c[:, 2]
    a = c[:, 3] * f(alpha)
    if rgb is not None:
        r, g, b = (np.full_like(r, f(x)) for x in rgb)
    else:
        mx, mn = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
        d = mx - mn
        with np.errstate(all="ignore"):
            h6 = np.where(mx == r, (g - b) / d, np.where(mx == g, (b - r) / d + f(2.0), (r - g) / d + f(4.0)))
            h6 = np.where(d > 0, h6, f(0.0)).astype(f)
            h = (h6 / f(6.0) + f(hsv[0])).astype(f)
            h = (h - np.floor(h)).astype(f)
            sat = np.clip(np.where(mx > 0, d / mx, f(0.0)).astype(f) * f(hsv[1]), 0, 1).astype(f)
        val = np.clip(mx * f(hsv[2]), 0, 1).astype(f)
        hh, vs = h * f(6.0), val * sat
        out = []
        for nn in (5.0, 3.0, 1.0):
            k = f(nn) + hh
            k = np.where(k >= 6, k - f(6.0), k).astype(f)
            t = np.maximum(f(0.0), np.minimum(np.minimum(k, f(4.0) - k), f(1.0)))
            out.append((val - vs * t).astype(f))
        r, g, b = out
    res = []
    for x in (r, g, b):
        x = ((x - f(0.5)) * (f(1.0) + f(contrast)) + f(0.5)).astype(f)
        x = (x * f(np.exp2(np.float64(exposure)))).astype(f)
        if gamma != 1.0:
            x = np.power(np.maximum(x, 0).astype(np.float64), gamma).astype(f)
        res.append(x)
    res.append(a)
    return np.stack([np.rint(np.clip(x, 0, 1) * f(255.0)).astype(np.uint8) for x in res], axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("kw,exact", [
    (dict(hsv=(0.0, 1.0, 1.0)), True),                                   # neutral: nothing changes
    (dict(hsv=(0.35, 0.6, 1.2), alpha=0.8), True),                       # hue rotation, desaturate, brighten
    (dict(hsv=(-0.2, 1.5, 0.7), contrast=0.5), True),                    # negative hue shift wraps; contrast
    (dict(rgb=(0.25, 1.0, 0.5), alpha=0.5, contrast=-0.3), True),        # override then contrast
    (dict(hsv=(0.1, 1.0, 1.0), exposure=0.75, gamma=2.2), False),        # exposure and gamma: powers, +-1 code
])
def test_basic_color_modifiers(sb, ob, ctx, kw, exact):
    """f4: the editor's BasicColorModifiers (HSV / contrast / exposure / gamma / alpha, or rgb override) on the selected
    Gaussians rewrite exactly the colour words the numpy statement predicts, leave the others alone, compose non-destructively
    (a second edit starts from the source again), restore brings the source back — and the frame is the oracle's frame of
    pods edited the same way."""
    import torch
    n, w, h = 20000, 640, 360
    g = sb.scenes.synthetic_gaussians(n, 37)
    pods = sb.pack_gaussians(g)
    stride = sb.pod_stride(0, 0)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    v.set_strict_exp(True)
    v.select_rect(150.0, 60.0, 500.0, 320.0)
    sel = v.read_selection()
    bits = ((sel[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
    assert 0 < bits.sum() < n
    v.apply_basic_color_modifiers(hsv=(0.5, 0.0, 0.3))  # an earlier edit that must leave no trace
    v.apply_basic_color_modifiers(**kw)
    t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(t, w, h)
    torch.cuda.synchronize()
    ptr, nbytes = v.device_pointers()["gaussians"]
    got = np.empty(nbytes, dtype=np.uint8)
    import ctypes
    torch.cuda.synchronize()
    cudart = ctypes.CDLL("libcudart.so.12")
    assert cudart.cudaMemcpy(ctypes.c_void_p(got.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), 2) == 0
    got = got.reshape(n, stride)
    src = pods.reshape(n, stride)
    exp_words = src[:, 12:16].copy()
    exp_words[bits] = _basic_color_np(src[bits, 12:16], **kw)
    if exact:
        assert np.array_equal(got[:, 12:16], exp_words)
    else:
        assert np.abs(got[:, 12:16].astype(np.int32) - exp_words.astype(np.int32)).max() <= 1
    assert np.array_equal(got[~bits], src[~bits]) and np.array_equal(np.delete(got, [12, 13, 14, 15], axis=1), np.delete(src, [12, 13, 14, 15], axis=1))
    oimg, _ = ob.render(ob.OracleModel(got.reshape(-1).copy(), n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
    assert np.array_equal(t.cpu().numpy(), oimg)
    v.restore_gaussians()
    torch.cuda.synchronize()
    back = np.empty(nbytes, dtype=np.uint8)
    assert cudart.cudaMemcpy(ctypes.c_void_p(back.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), 2) == 0
    assert np.array_equal(back, pods.reshape(-1).view(np.uint8))
    v.close()


def test_cpp_host_mirror_e2e(sb):
    """The typed C++ host mirror (host/splat_b200.hpp) runs the reference's viewer e2e test."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(sb.lib_path()), "..", "build", "e2e_viewer")
    assert os.path.exists(exe), "build/e2e_viewer missing: run __graft_entry__.build()"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout + out.stderr
