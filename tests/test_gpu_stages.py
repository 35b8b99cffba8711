"""The three stages as standalone objects on caller-owned buffers — Preprocessor<G, ()>, RadixSorter<()>,
Renderer<G, ()> (reference src/preprocessor.rs:370-450, src/radix_sorter.rs:71-96, src/renderer.rs:242-356) —
through the C ABI against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


class Buffers:
    """The buffers a reference user allocates themselves (src/lib.rs:100-199), as torch tensors."""

    def __init__(self, sb, pods, n, depth_factor=1):
        torch = _torch()
        self.gaussians = torch.from_numpy(pods).cuda()
        self.indirect_args = torch.tensor([6, 0, 0, 0], dtype=torch.int32, device="cuda")       # IndirectArgsBuffer::new
        self.sort_args = torch.tensor([1, 1, 1], dtype=torch.int32, device="cuda")              # RadixSortIndirectArgsBuffer::new
        self.indices = torch.zeros(max(n, 1), dtype=torch.int32, device="cuda")                 # IndirectIndicesBuffer
        # GaussiansDepthBuffer: the reference allocates 16 B per padded key; 4 are used
        self.depth = torch.full((max(sb.padded_key_count(n), 1) * depth_factor,), -1.0, dtype=torch.float32, device="cuda")


def scene(sb, n, seed, sh_fmt=0, cov_fmt=0):
    g = sb.scenes.synthetic_gaussians(n, seed)
    return g, sb.pack_gaussians(g, sh_fmt, cov_fmt)


def pods_of(sb, ob, cam_args, w, h, mode=0, mt=None):
    pos, yaw, pitch = cam_args
    cam, ocam = sb.camera_pod(pos, yaw, pitch, w, h), ob.camera_pod(pos, yaw, pitch, w, h)
    mtp = sb.model_transform_pod(*mt) if mt else sb.model_transform_pod()
    omt = ob.model_transform_pod(*mt) if mt else ob.model_transform_pod()
    return cam, ocam, mtp, omt, sb.gaussian_transform_pod(1.0, mode, 3, False, 3.0), ob.gaussian_transform_pod(1.0, mode, 3, False, 3.0)


@pytest.mark.parametrize("fmt", [(0, 0), (1, 1), (2, 2), (3, 0)])
@pytest.mark.parametrize("camera", ["outside", "inside"])
def test_preprocessor_standalone(sb, ob, ctx, fmt, camera):
    torch = _torch()
    n, w, h = 30011, 800, 450
    g, pods = scene(sb, n, 21, *fmt)
    cam_args = sb.scenes.CAMERA_OUTSIDE if camera == "outside" else sb.scenes.CAMERA_INSIDE
    cam, ocam, mtp, omt, gt, ogt = pods_of(sb, ob, cam_args, w, h)
    b = Buffers(sb, pods, n, depth_factor=4)
    pre = sb.Preprocessor(ctx, n, *fmt)
    bg = pre.create_bind_group(cam, mtp, gt, b.gaussians, b.indirect_args, b.sort_args, b.indices, b.depth)
    pre.preprocess(bg)
    torch.cuda.synchronize()
    o = ob.preprocess(ob.OracleModel(pods, n, *fmt, model_transform=omt), ocam, ogt)
    V = o["count"]
    assert b.indirect_args.cpu().numpy().astype(np.uint32).tolist() == list(o["draw_args"])
    assert b.sort_args.cpu().numpy().astype(np.uint32).tolist() == list(o["sort_args"])
    assert np.array_equal(b.indices.cpu().numpy().view(np.uint32)[:V], o["indices"][:V])  # ascending index order
    keys = b.depth.cpu().numpy()
    assert np.array_equal(keys[:V].view(np.uint32), o["keys"][:V].view(np.uint32))
    padded = int(o["sort_args"][0]) * 3840
    assert np.all(keys[V:padded] == np.float32(2.0)) and np.all(keys[padded:] == np.float32(-1.0))
    pre.close()


def test_preprocessor_selection_and_count(sb, ob, ctx):
    """bindings 8/9 and a gaussian_count below the buffer's length (preprocess(encoder, bind_group, gaussian_count))."""
    torch = _torch()
    n, w, h = 20000, 640, 360
    g, pods = scene(sb, n, 22)
    cam, ocam, mtp, omt, gt, ogt = pods_of(sb, ob, sb.scenes.CAMERA_OUTSIDE, w, h, mt=((1.0, -2.0, 0.5), (0.0, 0.38268343, 0.0, 0.92387953), (1.2, 0.9, 1.1)))
    rng = np.random.default_rng(5)
    words = rng.integers(0, 2**32, (n + 31) // 32, dtype=np.uint32)
    sel = torch.from_numpy(words.view(np.int32)).cuda()
    b = Buffers(sb, pods, n)
    pre = sb.Preprocessor(ctx, n)
    for invert in (0, 1):
        for count in (n, 12345, 0):
            bg = pre.create_bind_group(cam, mtp, gt, b.gaussians, b.indirect_args, b.sort_args, b.indices, b.depth, sel, invert)
            pre.preprocess(bg, count)
            torch.cuda.synchronize()
            stride = sb.pod_stride(0, 0)
            om = ob.OracleModel(pods[:count * stride], count, model_transform=omt, selection=words[:(count + 31) // 32], invert_selection=invert)
            o = ob.preprocess(om, ocam, ogt)
            V = o["count"]
            assert b.indirect_args.cpu().numpy().astype(np.uint32).tolist() == [6, V, 0, 0]
            assert b.sort_args.cpu().numpy().astype(np.uint32).tolist() == [(V + 3839) // 3840, 1, 1]
            assert np.array_equal(b.indices.cpu().numpy().view(np.uint32)[:V], o["indices"][:V])
            assert np.array_equal(b.depth.cpu().numpy()[:V].view(np.uint32), o["keys"][:V].view(np.uint32))
    pre.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_stage_chain_equals_viewer(sb, ob, ctx, mode):
    """Preprocessor -> RadixSorter -> Renderer on the caller's buffers produces the Viewer's frame, bit for bit, and the
    oracle's (strict exp)."""
    torch = _torch()
    n, w, h = 25000, 640, 360
    g, pods = scene(sb, n, 23)
    cam, ocam, mtp, omt, gt, ogt = pods_of(sb, ob, sb.scenes.CAMERA_OUTSIDE, w, h, mode=mode)
    b = Buffers(sb, pods, n)
    pre, sorter, ren = sb.Preprocessor(ctx, n), sb.RadixSorter(ctx, n), sb.Renderer(ctx, n)
    ren.set_strict_exp(True)
    pre.preprocess(pre.create_bind_group(cam, mtp, gt, b.gaussians, b.indirect_args, b.sort_args, b.indices, b.depth))
    sorter.sort(b.depth.data_ptr(), b.indices.data_ptr(), b.indirect_args.data_ptr() + 4, n)
    target = torch.full((h, w, 4), 77, dtype=torch.uint8, device="cuda")
    ren.render(target, w, h, ren.create_bind_group(cam, mtp, gt, b.gaussians, b.indices), b.indirect_args)
    torch.cuda.synchronize()

    v = sb.Viewer(ctx, pods, n)
    v.update_camera_with_pod(cam)
    v.update_gaussian_transform(1.0, mode, 3, False, 3.0)
    v.set_strict_exp(True)
    vt = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(vt, w, h)
    torch.cuda.synchronize()
    V = int(v.read_indirect_args()[0][1])
    assert int(b.indirect_args[1]) == V
    assert np.array_equal(b.indices.cpu().numpy().view(np.uint32)[:V], v.read_indices(V))
    assert np.array_equal(b.depth.cpu().numpy()[:V], v.read_depth_keys(V))
    assert torch.equal(target, vt)
    oimg, _ = ob.render(ob.OracleModel(pods, n, model_transform=omt), ocam, ogt, strict_exp=True)
    assert np.array_equal(target.cpu().numpy(), oimg)
    v.close(); pre.close(); sorter.close(); ren.close()


@pytest.mark.parametrize("fmt,target_format", [((0, 0), 0), ((1, 1), 1), ((2, 1), 3)])
def test_renderer_draws_callers_indices(sb, ob, ctx, fmt, target_format):
    """Renderer::render with indices no Preprocessor produced: arbitrary order, repeats, culled and behind-camera
    Gaussians (those are clipped as whole quads), an instance_count below the buffer length."""
    torch = _torch()
    n, w, h = 6000, 480, 270
    g, pods = scene(sb, n, 24, *fmt)
    cam, ocam, mtp, omt, gt, ogt = pods_of(sb, ob, sb.scenes.CAMERA_INSIDE, w, h)
    rng = np.random.default_rng(9)
    idx = rng.integers(0, n, n, dtype=np.uint32)  # with repeats, unsorted
    count = 4321
    gaussians = torch.from_numpy(pods).cuda()
    indices = torch.from_numpy(idx.view(np.int32)).cuda()
    args = torch.tensor([6, count, 0, 0], dtype=torch.int32, device="cuda")
    ren = sb.Renderer(ctx, n, *fmt, target_format=target_format)
    ren.set_strict_exp(True)
    dt = {0: torch.uint8, 1: torch.uint8, 3: torch.float32}[target_format]
    target = torch.zeros((h, w, 4), dtype=dt, device="cuda")
    bg = ren.create_bind_group(cam, mtp, gt, gaussians, indices)
    ren.render(target, w, h, bg, args)
    torch.cuda.synchronize()
    om = ob.OracleModel(pods, n, *fmt, model_transform=omt)
    oimg = np.zeros((h, w, 4), dtype=np.uint8 if target_format < 2 else np.float32)
    ob.draw(om, ocam, ogt, idx[:count], oimg, target_format=target_format, strict_exp=True)
    img = target.cpu().numpy()
    if target_format < 2:
        assert np.array_equal(img, oimg)
    else:
        assert np.abs(img - oimg).max() <= 1e-3  # float target tolerance (BASELINE north_star)
    assert img[..., :3].any()

    # render_with_pass: LoadOp::Load over the frame just drawn, with the next 1000 instances and a depth attachment
    depth = torch.full((h, w), 0.9990, dtype=torch.float32, device="cuda")
    indices2 = torch.from_numpy(idx[count:count + 1000].copy().view(np.int32)).cuda()
    pad = torch.zeros(n, dtype=torch.int32, device="cuda")
    pad[:1000] = indices2
    args2 = torch.tensor([6, 1000, 0, 0], dtype=torch.int32, device="cuda")
    ren.render(target, w, h, ren.create_bind_group(cam, mtp, gt, gaussians, pad), args2, depth=depth, compare=sb.COMPARE_LESS,
               depth_write=True, load_target=True)
    torch.cuda.synchronize()
    odepth = np.full((h, w), 0.9990, dtype=np.float32)
    ob.draw(om, ocam, ogt, idx[count:count + 1000], oimg, load=True, depth=odepth, compare=2, depth_write=True,
            target_format=target_format, strict_exp=True)
    img = target.cpu().numpy()
    if target_format < 2:
        assert np.array_equal(img, oimg)
    else:
        assert np.abs(img - oimg).max() <= 1e-3
    assert np.array_equal(depth.cpu().numpy(), odepth)
    ren.close()


def test_standalone_errors(sb, ctx):
    """TryFrom<wgpu::Buffer> size checks and construction errors surface as status codes, never aborts."""
    torch = _torch()
    n, w, h = 1000, 64, 64
    g, pods = scene(sb, n, 25)
    b = Buffers(sb, pods, n)
    cam = sb.camera_pod((0, 0, -30), 0.1, 0.1, w, h)
    mtp, gt = sb.model_transform_pod(), sb.gaussian_transform_pod()
    pre = sb.Preprocessor(ctx, n)
    with pytest.raises(sb.SplatError) as e:
        pre.preprocess(pre.create_bind_group(cam, mtp, gt, b.gaussians[:-16], b.indirect_args, b.sort_args, b.indices, b.depth))
    assert e.value.status == 5
    with pytest.raises(sb.SplatError) as e:
        pre.preprocess(pre.create_bind_group(cam, mtp, gt, b.gaussians, b.indirect_args, b.sort_args, b.indices[:-1], b.depth))
    assert e.value.status == 5
    with pytest.raises(sb.SplatError) as e:
        pre.preprocess(pre.create_bind_group(cam, mtp, gt, b.gaussians, b.indirect_args, b.sort_args, b.indices, b.depth[:-1]))
    assert e.value.status == 5
    with pytest.raises(sb.SplatError) as e:
        pre.preprocess(pre.create_bind_group(cam, mtp, gt, b.gaussians, b.indirect_args, b.sort_args, b.indices, b.depth), n + 1)
    assert e.value.status == 1
    pre.close()
    ctx.set_model_size_limit(1000)
    try:
        with pytest.raises(sb.SplatError) as e:
            sb.Preprocessor(ctx, n)
        assert e.value.status == 3  # ModelSizeExceedsDeviceLimit
    finally:
        ctx.set_model_size_limit(torch.cuda.mem_get_info()[0])  # the default: free device memory
    ren = sb.Renderer(ctx, n)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(sb.SplatError) as e:
        ren.render(target, w, h, ren.create_bind_group(cam, mtp, gt, b.gaussians[:-16], b.indices), b.indirect_args)
    assert e.value.status == 5
    with pytest.raises(sb.SplatError) as e:
        ren.render(target, w + 1, h, ren.create_bind_group(cam, mtp, gt, b.gaussians, b.indices), b.indirect_args)
    assert e.value.status == 1
    # an empty draw clears to BLACK
    ren.render(target.fill_(9), w, h, ren.create_bind_group(cam, mtp, gt, b.gaussians, b.indices), b.indirect_args)
    torch.cuda.synchronize()
    assert target[..., :3].eq(0).all() and target[..., 3].eq(255).all()
    ren.close()
