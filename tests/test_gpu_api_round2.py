"""GPU: the round-2 boundary additions and robustness fixes, through the C ABI, against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def idiff(a, b):
    return int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max())


def test_overflowing_frame_is_reported_by_the_next_enqueue(sb, ob, ctx):
    """A frame whose (splat, tile) duplicates exceed the capacity is flagged on the device and the NEXT enqueue returns
    SB_ERR_OVERFLOW without enqueuing (after growing the buffers): render() callers cannot miss it, and rendering again gives
    the complete frame."""
    import torch
    n, w, h = 400, 640, 360
    g = sb.scenes.synthetic_gaussians(n, 99, extent=1.5, log_scale=(-2.0, 0.5))  # near-camera splats covering every tile
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_INSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    v.set_strict_exp(True)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(target, w, h)
    torch.cuda.synchronize()
    need = v.read_frame_stats()["duplicates"]
    assert need > 8192
    v.reserve_duplicates(4096)  # smaller than the frame needs
    v.render(target, w, h)      # incomplete frame, flagged on the device
    torch.cuda.synchronize()
    assert v.read_frame_stats()["overflowed"]
    with pytest.raises(sb.SplatError) as e:
        v.render(target, w, h)
    assert e.value.status == 7 and "render again" in str(e.value)
    v.render(target, w, h)      # capacity was grown: complete again
    torch.cuda.synchronize()
    assert not v.read_frame_stats()["overflowed"]
    oimg, _ = ob.render(ob.OracleModel(pods, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
    assert idiff(target.cpu().numpy(), oimg) == 0
    v.close()


def test_duplicate_total_beyond_32_bits_is_flagged_not_wrapped(sb, ctx):
    """Many splats each covering the whole of a large target: the duplicate total exceeds 2^32.  The 64-bit scan reports
    it saturated (needed >= 2^32 - 1 > capacity) instead of wrapping to a small number that would hide the overflow."""
    import torch
    # axes clamp at 1024 px (utils.wesl:73-74): a splat covers at most 3072 x 3072 px = 192 x 192 = 36864 tiles;
    # 150000 x 36864 = 5.5e9 > 2^32
    n, w, h = 150_000, 4096, 4096
    g = np.zeros(n, dtype=sb.GAUSSIAN_DTYPE)
    g["pos"] = (0.0, 0.0, 0.3)
    g["pos"][:, 0] = np.linspace(-0.01, 0.01, n)
    g["rot"] = (0, 0, 0, 1)
    g["scale"] = (5.0, 5.0, 5.0)
    g["color"] = (200, 100, 50, 128)
    v = sb.Viewer(ctx, gaussians=g)
    v.update_camera((0, 0, 0), 0.0, 0.0, w, h)
    v.set_exact_cutoff(False)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(target, w, h)
    torch.cuda.synchronize()
    st = v.read_frame_stats()
    assert st["visible"] == n and st["overflowed"], st
    with pytest.raises(sb.SplatError) as e:
        v.render(target, w, h)
    assert e.value.status == 7 and ">= 4294967295" in str(e.value)
    v.close()


def test_render_batch_leaves_the_last_view_in_the_primary_viewer(sb, ob, ctx):
    """After sb_viewer_render_batch the viewer's public artefacts (indirect args, indices, keys) are those of the LAST view,
    for odd and for even batch sizes — what the reference's buffers hold after its last update_camera + render."""
    import torch
    n, w, h = 30000, 640, 360
    pods = sb.pack_gaussians(sb.scenes.synthetic_gaussians(n, 45))
    om = ob.OracleModel(pods, n)
    v = sb.Viewer(ctx, pods, n)
    stream = torch.cuda.Stream()
    for count in (4, 5, 1, 2):
        cam_args = [sb.scenes.orbit_camera(k, 9) for k in range(count)]
        cams = [sb.camera_pod(p, y, t, w, h) for p, y, t in cam_args]
        outs = [torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda") for _ in cams]
        v.render_batch(cams, targets=outs, width=w, height=h, stream=stream)
        stream.synchronize()
        p, y, t = cam_args[-1]
        pre = ob.preprocess(om, ob.camera_pod(p, y, t, w, h), ob.gaussian_transform_pod())
        draw, _ = v.read_indirect_args(stream)
        assert int(draw[1]) == pre["count"], f"batch of {count}"
        ok, oi = ob.radix_sort(pre["keys"][: pre["count"]].view(np.uint32), pre["indices"][: pre["count"]])
        assert np.array_equal(v.read_indices(pre["count"], stream), oi)
        assert np.array_equal(v.read_depth_keys(pre["count"], stream).view(np.uint32), ok)
        # the stage-wise draw on the primary viewer reproduces the last view
        again = torch.zeros_like(outs[-1])
        v.draw(again, w, h, stream=stream)
        stream.synchronize()
        assert torch.equal(again, outs[-1])
    v.close()


def test_renderer_accepts_unaligned_index_subbuffer_and_clamps_the_count(sb, ob, ctx):
    """Renderer<G, ()> on a caller's buffers: an index list that is only 4-byte aligned (a slice of a larger buffer) and an
    instance_count larger than the model are handled like WebGPU's robust buffer access — no fault, extra instances draw
    nothing."""
    import torch
    n, w, h = 5000, 480, 270
    pods_np = sb.pack_gaussians(sb.scenes.synthetic_gaussians(n, 46))
    pods = torch.from_numpy(pods_np).cuda()
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    cam, mt, gt = sb.camera_pod(pos, yaw, pitch, w, h), sb.model_transform_pod(), sb.gaussian_transform_pod()
    rng = np.random.default_rng(3)
    order = rng.permutation(n).astype(np.uint32)
    backing = torch.zeros(n + 8, dtype=torch.int32, device="cuda")
    sub = backing[1: 1 + n]  # 4-byte aligned, not 16
    assert sub.data_ptr() % 16 == 4
    sub.copy_(torch.from_numpy(order.view(np.int32)).cuda())
    r = sb.Renderer(ctx, n)
    r.set_strict_exp(True)
    bg = sb.Renderer.create_bind_group(cam, mt, gt, pods, sub)
    om, ocam, ogt = ob.OracleModel(pods_np, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod()
    for count in (n, n - 123, n + 1000):
        args = torch.tensor([6, count, 0, 0], dtype=torch.int32, device="cuda")
        t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        r.render(t, w, h, bg, args)
        torch.cuda.synchronize()
        want = np.zeros((h, w, 4), dtype=np.uint8)
        ob.draw(om, ocam, ogt, order[: min(count, n)], want, strict_exp=True)
        assert idiff(t.cpu().numpy(), want) == 0, f"instance_count {count}"
    r.close()


def test_multi_model_non_pod_updates_device_models_and_depth_pass(sb, ob, ctx):
    """MultiModelViewer: update_camera / update_model_transform / update_gaussian_transform (non-pod,
    src/multi_model.rs:398-464), insert_model from source Gaussians and on an adopted device buffer, and the models drawn
    inside a caller's pass with a depth attachment (new_with_options(depth_stencil) + renderer.render_with_pass)."""
    import torch
    n, w, h = 8000, 512, 288
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    mm = sb.MultiModelViewer(ctx)
    mm.update_camera(pos, yaw, pitch, w, h)
    mm.update_gaussian_transform(1.0, sb.MODE_SPLAT, 2, False, 2.5)
    ogt = ob.gaussian_transform_pod(1.0, ob.MODE_SPLAT, 2, False, 2.5)
    ocam = ob.camera_pod(pos, yaw, pitch, w, h)
    gs = [sb.scenes.synthetic_gaussians(n, 300 + k, extent=6.0, log_scale=(-3.5, -2.0)) for k in range(3)]
    pods = [sb.pack_gaussians(g) for g in gs]
    dev = torch.from_numpy(pods[2]).cuda()
    assert mm.insert_model(10, pods[0], n) is False
    assert mm.insert_model_from_gaussians(11, gs[1]) is False
    assert mm.insert_model_from_device(12, dev, n) is False
    mts = {10: ((-5.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0)),
           11: ((0.0, 1.0, 0.0), (0.0, float(np.sin(0.2)), 0.0, float(np.cos(0.2))), (1.2, 1.2, 1.2)),
           12: ((5.0, 0.0, 2.0), (float(np.sin(0.1)), 0.0, 0.0, float(np.cos(0.1))), (0.9, 1.0, 1.1))}
    for k, mt in mts.items():
        mm.update_model_transform(k, *mt)
    with pytest.raises(sb.SplatError) as e:
        mm.update_model_transform(99, *mts[10])
    assert e.value.status == 4
    om = {k: ob.OracleModel(pods[k - 10], n, model_transform=ob.model_transform_pod(*mts[k])) for k in mts}
    order = [12, 10, 11]
    t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    mm.render(t, w, h, order)
    torch.cuda.synchronize()
    oimg, _ = ob.render([om[k] for k in order], ocam, ogt)
    assert idiff(t.cpu().numpy(), oimg) <= 2
    # caller's pass: existing colour + a depth wall; models composite in key order, each depth-tested and writing depth
    rng = np.random.default_rng(6)
    colour = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    colour[..., 3] = 255
    zz = np.linspace(0.9955, 0.9985, w, dtype=np.float32)[None, :].repeat(h, 0).copy()
    for compare, write in ((sb.COMPARE_LESS, True), (sb.COMPARE_GREATER_EQUAL, False)):
        tc, td = torch.from_numpy(colour.copy()).cuda(), torch.from_numpy(zz.copy()).cuda()
        mm.render_with_pass(tc, w, h, order, depth=td, compare=compare, depth_write=write, load_target=True)
        torch.cuda.synchronize()
        ocol, oz = colour.copy(), zz.copy()
        for k in order:
            ob.render_pass(om[k], ocam, ogt, ocol, True, depth=oz, compare=compare, depth_write=write)
        assert idiff(tc.cpu().numpy(), ocol) <= 2, (compare, write)
        gz = td.cpu().numpy()
        # depth writes depend only on the discard test and the compare: exact
        assert np.array_equal(gz, oz), "depth attachment differs"
        assert (gz != zz).any() == write
    mm.close()
