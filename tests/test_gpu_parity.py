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
    bpp = {0: (torch.uint8, 4), 1: (torch.uint8, 4), 2: (torch.float16, 4), 3: (torch.float32, 4), 4: (torch.uint8, 4), 5: (torch.uint8, 4)}[target_format]
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


def img_diff(r):
    a, b = r["img"], r["oimg"]
    if a.dtype == np.uint8:
        return np.abs(a.astype(np.int32) - b.astype(np.int32)).max()
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


# every pod format the reference can instantiate (SH x cov3d): strides pinned in test_abi.py
FORMATS = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (3, 2)]


@pytest.mark.parametrize("sh_fmt,cov_fmt", FORMATS)
def test_pod_formats(sb, ob, ctx, sh_fmt, cov_fmt):
    n = 6000 + 37 * sh_fmt + 11 * cov_fmt  # ragged: not a multiple of any tile size
    g, pods = make_scene(sb, ob, n, 10 + sh_fmt * 3 + cov_fmt, sh_fmt, cov_fmt)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_INSIDE, 480, 270, sh_fmt, cov_fmt)
    check_artifacts(ob, r, n)
    assert img_diff(r) <= 1, f"max-abs {img_diff(r)}"


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("target_format", [0, 1, 2, 3, 4, 5])
def test_modes_and_targets(sb, ob, ctx, mode, target_format):
    n = 8000
    g, pods = make_scene(sb, ob, n, 40 + mode)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, 512, 288, mode=mode, target_format=target_format)
    check_artifacts(ob, r, n)
    d = img_diff(r)
    if target_format in (0, 1):
        assert d <= 1, f"unorm8 max-abs {d}/255"       # tolerance 2/255 (north star); strict exp => expect 0-1
        assert np.all(r["img"][..., 3] == 255)
    elif target_format in (4, 5):
        # *Srgb attachments: decode - blend - encode per blend through the shared exact tables: strict exp => identical codes
        assert d == 0, f"sRGB max-abs {d}/255"
        assert np.all(r["img"][..., 3] == 255) and r["img"][..., :3].max() > 0
    else:
        assert d <= 1e-3, f"float target max-abs {d}"  # north-star tolerance for float targets


def test_srgb_target_semantics(sb, ob, ctx):
    """An sRGB attachment is not a relabelled unorm8 one: the same scene rendered to Rgba8UnormSrgb holds the sRGB ENCODING of
    (roughly) what the linear target holds, fast-exp frames stay within tolerance of the oracle, bgra swizzles, and compositing
    over existing content (LoadOp::Load) decodes what it finds."""
    torch = _torch()
    n, w, h = 12000, 512, 288
    g, pods = make_scene(sb, ob, n, 83)
    lin = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, w, h, target_format=0, strict=False)
    srgb = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, w, h, target_format=4, strict=False)
    bgra = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, w, h, target_format=5, strict=False)
    assert img_diff(srgb) <= 2 and img_diff(bgra) <= 2
    assert np.array_equal(srgb["img"][..., [2, 1, 0, 3]], bgra["img"])
    x = lin["img"][..., :3].astype(np.float64) / 255.0
    enc = np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(x, 1 / 2.4) - 0.055) * 255.0
    lit = lin["img"][..., :3] > 24  # away from black, where one linear code spans many sRGB codes
    assert np.abs(enc - srgb["img"][..., :3].astype(np.float64))[lit].mean() < 6.0
    assert srgb["img"][..., :3].astype(np.int64).sum() > 1.5 * lin["img"][..., :3].astype(np.int64).sum()
    # LoadOp::Load over a mid-grey sRGB background
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n, target_format=4)
    v.set_strict_exp(True)
    v.update_camera_with_pod(sb.camera_pod(pos, yaw, pitch, w, h))
    bg = np.zeros((h, w, 4), dtype=np.uint8)
    bg[..., 0], bg[..., 1], bg[..., 2], bg[..., 3] = 188, 128, 64, 255
    target = torch.from_numpy(bg.copy()).cuda()
    v.render_with_pass(target, w, h, load_target=True)
    torch.cuda.synchronize()
    got = target.cpu().numpy()
    v.close()
    om = ob.OracleModel(pods, n)
    exp = ob.render_pass(om, ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), bg.copy(), True, target_format=4, strict_exp=True)
    assert np.array_equal(got, exp)
    assert np.array_equal(got[0, 0], bg[0, 0]) or got[..., :3].max() > 0  # untouched pixels keep their codes


@pytest.mark.parametrize("sh_deg,no_sh0,std_dev,size", [(0, False, 3.0, 1.0), (1, False, 2.0, 1.0), (2, True, 3.0, 0.5),
                                                        (3, True, 1.0, 2.0), (3, False, 0.0, 1.0)])
def test_gaussian_transform_knobs(sb, ob, ctx, sh_deg, no_sh0, std_dev, size):
    n = 5000
    g, pods = make_scene(sb, ob, n, 60 + sh_deg)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_INSIDE, 400, 300, sh_deg=sh_deg, no_sh0=no_sh0, std_dev=std_dev, size=size)
    check_artifacts(ob, r, n)
    assert img_diff(r) <= 1


def test_model_transform(sb, ob, ctx):
    n = 7000
    g, pods = make_scene(sb, ob, n, 70)
    q = np.array([0.1, 0.7, -0.2, 0.6], dtype=np.float32)
    q /= np.linalg.norm(q)
    mt = ((1.5, -2.0, 3.0), tuple(q), (1.3, 0.8, 1.1))
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, 640, 360, model_transform=mt)
    check_artifacts(ob, r, n)
    assert img_diff(r) <= 1


@pytest.mark.parametrize("invert", [1, 0])
def test_selection_mask(sb, ob, ctx, invert):
    n = 9000
    g, pods = make_scene(sb, ob, n, 80)
    rng = np.random.default_rng(5)
    words = rng.integers(0, 2**32, size=(n + 31) // 32, dtype=np.uint64).astype(np.uint32)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, 480, 270, selection=words, invert=invert)
    check_artifacts(ob, r, n)
    assert 0 < r["V"] < n
    assert img_diff(r) <= 1


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 383, 384, 385, 3839, 3840, 3841])
def test_ragged_sizes(sb, ob, ctx, n):
    g = sb.scenes.synthetic_gaussians(max(n, 1), 90 + n, extent=3.0, log_scale=(-3.0, -1.5))[:n]
    pods = sb.pack_gaussians(g) if n else np.zeros(0, dtype=np.uint8)
    r = run_frame(sb, ob, ctx, pods, n, ((0.0, 0.0, -8.0), 0.05, 0.05), 320, 200)
    check_artifacts(ob, r, n)
    assert img_diff(r) <= 1


def test_huge_near_splats(sb, ob, ctx):
    """Axes clamp at 1024 px (utils.wesl:73-74): near-camera splats cover every tile."""
    n = 300
    g = sb.scenes.synthetic_gaussians(n, 99, extent=1.5, log_scale=(-2.0, 0.5))
    pods = sb.pack_gaussians(g)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_INSIDE, 640, 360)
    check_artifacts(ob, r, n)
    assert not r["stats"]["overflowed"]
    assert img_diff(r) <= 1


def test_deep_low_alpha_stack(sb, ob, ctx):
    """Per-blend re-quantisation on unorm8: many faint layers over one pixel region."""
    n = 4000
    g = sb.scenes.synthetic_gaussians(n, 123, extent=0.2, log_scale=(-2.5, -2.0))
    g["pos"][:, 2] = np.linspace(2.0, 6.0, n)
    g["color"][:, 3] = 3  # alpha ~ 0.012
    pods = sb.pack_gaussians(g)
    r = run_frame(sb, ob, ctx, pods, n, ((0.0, 0.0, 0.0), 0.0, 0.0), 256, 256)
    check_artifacts(ob, r, n)
    assert img_diff(r) == 0
    # a float-accumulating compositor would differ from the per-blend quantised result by many levels
    fr = run_frame(sb, ob, ctx, pods, n, ((0.0, 0.0, 0.0), 0.0, 0.0), 256, 256, target_format=3)
    f8 = np.clip(np.rint(fr["img"][..., :3] * 255.0), 0, 255).astype(np.int32)
    assert np.abs(f8 - r["img"][..., :3].astype(np.int32)).max() > 3


@pytest.mark.parametrize("target_format", [0, 3])
def test_overbright_sh_colours(sb, ob, ctx, target_format):
    """SH lobes push view_color far above 1 (utils.wesl:82-135 clamps only below).  A fixed-point target clamps the
    SOURCE colour to [0,1] before the blend (Vulkan 1.3 spec 29.1); a float target keeps the over-bright value."""
    n = 5000
    g = sb.scenes.synthetic_gaussians(n, 321)
    g["sh"] *= 12.0
    pods = sb.pack_gaussians(g)
    r = run_frame(sb, ob, ctx, pods, n, sb.scenes.CAMERA_OUTSIDE, 480, 270, target_format=target_format)
    check_artifacts(ob, r, n)
    if target_format == 0:
        assert img_diff(r) == 0
    else:
        assert r["img"][..., :3].max() > 1.5, "float target must keep over-bright colours"
        d = np.abs(r["img"].astype(np.float64) - r["oimg"].astype(np.float64))
        assert (d <= 1e-3 * np.maximum(1.0, np.abs(r["oimg"]))).all()


def test_equal_depth_ties_keep_index_order(sb, ob, ctx):
    """equal keys are the norm; the canonical order inside a run is ascending index."""
    n = 5000
    g = sb.scenes.synthetic_gaussians(n, 77)
    g["pos"][:, 2] = np.repeat(np.linspace(-5, 5, 10), n // 10).astype(np.float32)
    pods = sb.pack_gaussians(g)
    r = run_frame(sb, ob, ctx, pods, n, ((0.0, 0.0, -30.0), 0.0, 0.0), 480, 270)
    check_artifacts(ob, r, n)
    k = r["keys"][: r["V"]].view(np.uint32)
    assert len(np.unique(k)) < r["V"] // 4
    same = k[1:] == k[:-1]
    assert np.all(r["idx"][1:][same] > r["idx"][:-1][same])
    assert img_diff(r) == 0


def test_screen_strips_match_full_frame(sb, ob, ctx):
    """Config 5b: a frame rendered as horizontal strips equals the full frame bit for bit."""
    torch = _torch()
    n, w, h = 20000, 640, 360
    g, pods = make_scene(sb, ob, n, 31)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    cam = sb.camera_pod(pos, yaw, pitch, w, h)
    v = sb.Viewer(ctx, pods, n)
    v.update_camera_with_pod(cam)
    full = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(full, w, h)
    for parts in (2, 3, 8):
        out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        bounds = [h * i // parts for i in range(parts + 1)]
        for i in range(parts):
            r0, r1 = bounds[i], bounds[i + 1]
            strip = torch.zeros((r1 - r0, w, 4), dtype=torch.uint8, device="cuda")
            v.render(strip, w, h, row0=r0, rows=r1 - r0)
            out[r0:r1] = strip
        torch.cuda.synchronize()
        assert torch.equal(out, full), f"{parts} strips differ from the full frame"
    ocam = ob.camera_pod(pos, yaw, pitch, w, h)
    om = ob.OracleModel(pods, n)
    ostrip, _ = ob.render(om, ocam, ob.gaussian_transform_pod(), row0=100, rows=57)
    assert np.abs(full[100:157].cpu().numpy().astype(np.int32) - ostrip.astype(np.int32)).max() <= 2
    v.close()


@pytest.mark.parametrize("cam_name,mode", [("outside", 0), ("inside", 0), ("inside", 1), ("outside", 2)])
def test_strip_cull_subsets_reassemble_the_frame(sb, ob, ctx, cam_name, mode):
    """Config 5b as the north star states it: every strip culls to and sorts its OWN visible set.  The strips still reassemble
    the full frame bit for bit; a strip's sorted list is a subsequence of the full frame's (same relative order, same keys);
    together the strips cover every splat that has a tile at all, and a strip's list is a fraction of the full one."""
    torch = _torch()
    n, w, h = 60000, 640, 368
    g, pods = make_scene(sb, ob, n, 77)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE if cam_name == "outside" else sb.scenes.CAMERA_INSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera_with_pod(sb.camera_pod(pos, yaw, pitch, w, h))
    v.update_gaussian_transform(1.0, mode, 3, False, 3.0)
    full = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(full, w, h)
    torch.cuda.synchronize()
    V = int(v.read_indirect_args()[0][1])
    fidx, fkeys = v.read_indices(V), v.read_depth_keys(V).view(np.uint32)
    pos_of = np.full(n, -1, dtype=np.int64)
    pos_of[fidx] = np.arange(V)
    dup_full = v.read_frame_stats()["duplicates"]
    v.set_strip_cull(True)
    for parts in (2, 5):
        out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        covered = np.zeros(n, dtype=bool)
        total, dups = 0, 0
        for i in range(parts):
            r0, rows = sb.sharding.strip_rows(h, parts, i)
            v.render(out[r0:r0 + rows], w, h, row0=r0, rows=rows)
            torch.cuda.synchronize()
            draw, disp = v.read_indirect_args()
            Vs = int(draw[1])
            assert list(draw) == [6, Vs, 0, 0] and list(disp) == [(Vs + 3839) // 3840, 1, 1]
            sidx, skeys = v.read_indices(Vs), v.read_depth_keys(Vs).view(np.uint32)
            where = pos_of[sidx]
            assert np.all(where >= 0), "a strip's splat is not in the full frame's visible set"
            assert np.all(np.diff(where) > 0), "a strip's order is not the full frame's order"
            assert np.array_equal(skeys, fkeys[where])
            covered[sidx] = True
            total += Vs
            dups += v.read_frame_stats()["duplicates"]
        assert torch.equal(out, full), f"{parts} culled strips differ from the full frame"
        assert dups == dup_full, "the strips' (splat, tile) duplicates must partition the full frame's"
        assert total < V * (1.0 + 0.9 * (parts - 1) / parts) or V < 1000, "strip lists are not smaller than the full list"
    v.set_strip_cull(False)
    v.render(full, w, h)  # the option is per render: full-frame artefacts again
    torch.cuda.synchronize()
    assert int(v.read_indirect_args()[0][1]) == V
    v.close()


def _strip_worker(rank, world, port, n, w, h, q):
    import os
    import sys
    import numpy as np
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "wgpu-3dgs-viewer_b200"))
    import splat_b200 as sb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = sb.Context(rank)
    pods = sb.pack_gaussians(sb.scenes.synthetic_gaussians(n, 91))
    v = sb.Viewer(ctx, pods, n)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v.update_camera_with_pod(sb.camera_pod(pos, yaw, pitch, w, h))
    res = {}
    sel = np.random.default_rng(4).integers(0, 2**32, size=(n + 31) // 32, dtype=np.uint64).astype(np.uint32)
    for mode in ("peer", "sendrecv", "exchange", "exchange+selection"):
        if mode == "exchange+selection":  # the selection mask is indexed by the Gaussian's index in the WHOLE model on every slice
            v.enable_selection(True)
            v.set_selection(sel)
        sf = sb.sharding.StripFrame(ctx, v, w, h, 4, world, rank, dst=0, mode="peer" if mode.startswith("exchange") else mode,
                                    balance=mode.startswith("exchange"), partition_cull=mode.startswith("exchange"))
        for _ in range(3):
            frame = sf.render()
        torch.cuda.synchronize()
        dist.barrier()
        lists = None
        if mode.startswith("exchange"):
            # this rank's artefacts describe its strip's visible set exactly as the strip-culling K1 leaves them
            Vx = int(v.read_indirect_args()[0][1])
            ix, kx = v.read_indices(Vx), v.read_depth_keys(Vx).view(np.uint32)
            r0, rows = sf.bounds[rank]
            chk = torch.zeros((max(rows, 1), w, 4), dtype=torch.uint8, device="cuda")
            v.set_strip_cull(True)
            v.render(chk, w, h, row0=r0, rows=rows)
            torch.cuda.synchronize()
            Vs = int(v.read_indirect_args()[0][1])
            lists = bool(Vx == Vs and np.array_equal(ix, v.read_indices(Vs)) and np.array_equal(kx, v.read_depth_keys(Vs).view(np.uint32)))
            agree = torch.tensor([int(lists)], device="cuda")
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            lists = bool(agree.item())
        if rank == 0:
            got = frame.clone()
            v.set_strip_cull(False)
            ref = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
            v.render(ref, w, h)
            torch.cuda.synchronize()
            res[mode] = (sf.mode, bool(torch.equal(got, ref)), lists)
        dist.barrier()
        sf.close()
    if rank == 0:
        q.put(res)
    v.close()
    ctx.close()
    dist.destroy_process_group()


def test_strips_over_two_gpus_land_in_one_frame(sb):
    """Two processes, two GPUs, NCCL: each rank renders its own strip (own cull, own sort) into the owner's frame — once through
    the peer-mapped frame (the rasterizer's stores cross NVLink), once through the batched send/recv — and the result equals the
    single-GPU frame.  Skipped on a one-GPU box."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_strip_worker, args=(r, 2, port, 200_000, 1920, 1080, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["sendrecv"][:2] == ("sendrecv", True)
    assert res["peer"][1], "strips written through the peer-mapped frame differ from the single-GPU frame"
    for mode in ("exchange", "exchange+selection"):
        assert res[mode][1], f"{mode}: partitioned-cull strips differ from the single-GPU frame"
        assert res[mode][2], f"{mode}: a rank's exchanged visible list differs from the strip-culling Preprocessor's"


@pytest.mark.parametrize("explicit_stream", [False, True])
def test_multi_model_draw_order_and_selection(sb, ob, ctx, explicit_stream):
    """Config 4: models composited in caller key order (multi_model.rs:505-527), per-model
    transforms, per-model selection masks — on the legacy default stream and on a stream of the caller's (the models'
    preprocess / sort / binning chains run on internal side streams, the rasters in key order on the caller's)."""
    torch = _torch()
    stream = torch.cuda.Stream() if explicit_stream else None
    w, h, n = 640, 360, 6000
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    cam = sb.camera_pod(pos, yaw, pitch, w, h)
    ocam = ob.camera_pod(pos, yaw, pitch, w, h)
    mm = sb.MultiModelViewer(ctx)
    mm.update_camera_with_pod(cam)
    gt = sb.gaussian_transform_pod()
    mm.update_gaussian_transform_with_pod(gt)
    omodels = {}
    rng = np.random.default_rng(3)
    for k in range(4):
        g = sb.scenes.synthetic_gaussians(n, 200 + k, extent=6.0, log_scale=(-3.5, -2.0))
        pods = sb.pack_gaussians(g)
        assert mm.insert_model(k, pods, n) is False
        a = 0.3 * k
        mt = ((6.0 * (k - 1.5), 0.0, 0.0), (0.0, float(np.sin(a / 2)), 0.0, float(np.cos(a / 2))), (1 + 0.05 * k,) * 3)
        mm.update_model_transform_with_pod(k, sb.model_transform_pod(*mt))
        sel = None
        if k % 2 == 1:
            sel = rng.integers(0, 2**32, size=(n + 31) // 32, dtype=np.uint64).astype(np.uint32)
            mm.set_selection(k, sel, invert=(k == 1))
        omodels[k] = ob.OracleModel(pods, n, model_transform=ob.model_transform_pod(*mt), selection=sel, invert_selection=int(k == 1))
    for order in ([0, 1, 2, 3], [3, 1, 0, 2], [2], [1, 1, 0]):
        target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        mm.render(target, w, h, order, stream=stream)
        mm.render(target, w, h, order, stream=stream)  # back to back: the second frame reuses every buffer of the first
        torch.cuda.synchronize()
        oimg, _ = ob.render([omodels[k] for k in order], ocam, ob.gaussian_transform_pod())
        d = np.abs(target.cpu().numpy().astype(np.int32) - oimg.astype(np.int32)).max()
        assert d <= 2, f"order {order}: max-abs {d}"
        for k in set(order):
            pre = ob.preprocess(omodels[k], ocam, ob.gaussian_transform_pod())
            idx, V = mm.read_model_indices(k, pre["count"])
            assert V == pre["count"]
            _, oi = ob.radix_sort(pre["keys"][:V].view(np.uint32), pre["indices"][:V])
            assert np.array_equal(idx, oi)
    with pytest.raises(sb.SplatError) as e:
        mm.render(target, w, h, [0, 9])
    assert e.value.status == 4  # ModelNotFound
    assert mm.remove_model(2) is True and mm.remove_model(2) is False
    with pytest.raises(sb.SplatError):
        mm.render(target, w, h, [2])
    mm.render(target, w, h, [])  # a pass that only clears
    torch.cuda.synchronize()
    t = target.cpu().numpy()
    assert np.all(t[..., :3] == 0) and np.all(t[..., 3] == 255)
    mm.close()


def test_viewport_rect_selection(sb, ob, ctx):
    """Next-row f1: selection::viewport evaluation with an analytic rectangle mask."""
    n, w, h = 30000, 640, 360
    g, pods = make_scene(sb, ob, n, 55)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    v.select_rect(160.0, 90.0, 480.0, 270.0)
    words = v.read_selection()
    om = ob.OracleModel(pods, n)
    ow = ob.select_rect(om, ob.camera_pod(pos, yaw, pitch, w, h), 160.0, 90.0, 480.0, 270.0)
    assert np.array_equal(words, ow)
    assert 0 < np.unpackbits(words.view(np.uint8)).sum() < n
    v.close()


def test_viewport_brush_selection(sb, ob, ctx):
    """selection::viewport with the brush mask (viewport_texture_brush.wesl: capsules along the stroke), strokes
    accumulating like in the reference's mask texture; the mask then drives the preprocess cull."""
    n, w, h = 30000, 800, 450
    g, pods = make_scene(sb, ob, n, 58)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    v.enable_selection(True)
    om = ob.OracleModel(pods, n)
    ocam = ob.camera_pod(pos, yaw, pitch, w, h)
    stroke1 = [(100.0, 80.0), (380.5, 200.25), (420.0, 400.0)]
    stroke2 = [(600.0, 120.0)]  # a dab
    v.select_brush(stroke1, 37.5)
    m1 = v.read_selection()
    o1 = ob.select_brush(om, ocam, stroke1, 37.5)
    assert np.array_equal(m1, o1) and m1.any()
    v.select_brush(stroke2, 60.0, accumulate=True)
    m2 = v.read_selection()
    o2 = ob.select_brush(om, ocam, stroke2, 60.0, accumulate=True, dest=o1)
    assert np.array_equal(m2, o2) and (m2 & ~m1).any()
    # zero radius: only texel centres exactly on the stroke; degenerate segment
    v.select_brush([(10.5, 10.5), (10.5, 10.5)], 0.0)
    assert np.array_equal(v.read_selection(), ob.select_brush(om, ocam, [(10.5, 10.5), (10.5, 10.5)], 0.0))
    with pytest.raises(sb.SplatError):
        v.select_brush(np.zeros((65, 2)), 1.0)
    v.close()


@pytest.mark.parametrize("compare,write", [(2, True), (4, False), (5, True), (8, False), (1, True)])
@pytest.mark.parametrize("target_format", [0, 3])
def test_render_with_pass_depth_stencil(sb, ob, ctx, compare, write, target_format):
    """Renderer::render_with_pass inside a caller's pass (src/renderer.rs:187-195) with the pipeline's depth_stencil state
    (:123, :304): LoadOp::Load of an existing colour target, fragments depth-tested with the splat's centre z against a
    Depth32Float attachment the caller filled, depth written back when depth_write_enabled."""
    torch = _torch()
    n, w, h = 12000, 512, 288
    g, pods = make_scene(sb, ob, n, 140 + compare)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    rng = np.random.default_rng(5)
    # caller's pass so far: a colour gradient and a depth "wall" cutting through the cloud (ndc z of the cloud ~ 0.9966..0.9975)
    if target_format == 0:
        colour = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        colour[..., 3] = 255
    else:
        colour = rng.random(size=(h, w, 4), dtype=np.float32)
        colour[..., 3] = 1.0
    zz = np.linspace(0.9962, 0.9978, w, dtype=np.float32)[None, :].repeat(h, 0).copy()
    zz[::7] = 1.0  # some rows wide open
    v = sb.Viewer(ctx, pods, n, target_format=target_format)
    v.update_camera(pos, yaw, pitch, w, h)
    v.set_strict_exp(True)
    t = torch.from_numpy(colour.copy()).cuda()
    d = torch.from_numpy(zz.copy()).cuda()
    v.render_with_pass(t, w, h, depth=d, compare=compare, depth_write=write, load_target=True)
    torch.cuda.synchronize()
    om = ob.OracleModel(pods, n)
    ocol, oz = colour.copy(), zz.copy()
    ob.render_pass(om, ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), ocol, True, depth=oz, compare=compare,
                   depth_write=write, target_format=target_format, strict_exp=True)
    got, gz = t.cpu().numpy(), d.cpu().numpy()
    if target_format == 0:
        assert np.array_equal(got, ocol)
    else:
        assert np.abs(got - ocol).max() <= 1e-3
    assert np.array_equal(gz, oz), "depth attachment differs"
    if compare == 1:
        assert np.array_equal(got, colour) and np.array_equal(gz, zz)  # Never: nothing drawn
    elif write:
        assert (gz != zz).any()
    else:
        assert np.array_equal(gz, zz)
    if compare in (2, 5):
        assert (got != colour).any()
    # without a depth attachment and with LoadOp::Clear the pass equals Viewer::render
    t2, t3 = torch.zeros_like(t), torch.zeros_like(t)
    v.render_with_pass(t2, w, h, load_target=False)
    v.render(t3, w, h)
    torch.cuda.synchronize()
    assert torch.equal(t2, t3)
    v.close()


def test_standalone_radix_sorter(sb, ctx):
    """RadixSorter<()>: stable ascending (key, payload) sort with a device-side count."""
    torch = _torch()
    rng = np.random.default_rng(11)
    for n, bits in ((1, 32), (4095, 32), (4096, 32), (100_001, 32), (250_000, 13), (70_000, 17)):
        keys = rng.integers(0, 2**bits, size=n, dtype=np.uint64).astype(np.uint32)
        if bits == 32:
            keys[: n // 2] &= np.uint32(0xFF00FFFF)  # plenty of ties
        vals = np.arange(n, dtype=np.uint32)
        dk = torch.from_numpy(keys.view(np.int32)).cuda()
        dv = torch.from_numpy(vals.view(np.int32)).cuda()
        cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
        s = sb.RadixSorter(ctx, n + 100)
        s.sort(dk.data_ptr(), dv.data_ptr(), cnt.data_ptr(), n, 0, bits)
        torch.cuda.synchronize()
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(dk.cpu().numpy().view(np.uint32), keys[order])
        assert np.array_equal(dv.cpu().numpy().view(np.uint32), vals[order])
        s.close()


@pytest.mark.parametrize("impl", ["v1", "v3", "v4"])
def test_radix_sorter_both_pass_kernels(sb, impl):
    """Both digit-pass kernels (SB_SORT_IMPL=v1: 8192-pair tiles, shared-memory peer masks; v3: 4096-pair tiles, vote ranking,
    cp.async payload) forced on every shape: odd bit ranges, begin_bit > 0, a device count below the capacity, an empty
    input, constant keys (every pass is the identity and is skipped) and sorted / reversed inputs.  Subprocess: the choice is
    read once per process."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import splat_b200 as sb
ctx = sb.Context(0)
rng = np.random.default_rng(23)
cases = [(1, 0, 32, 1), (4097, 0, 32, 4097), (12289, 0, 17, 12289), (50000, 0, 20, 50000), (50000, 0, 25, 50000),
         (65536, 4, 21, 65536), (300000, 0, 32, 123457), (5000, 0, 32, 0), (40000, 8, 16, 40000), (33333, 0, 1, 33333)]
for n, b0, b1, count in cases:
    for kind in ("random", "constant", "sorted", "reversed"):
        if kind == "random":
            keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        elif kind == "constant":
            keys = np.full(n, 0xDEADBEEF, dtype=np.uint32)
        else:
            keys = np.sort(rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32))
            if kind == "reversed":
                keys = keys[::-1].copy()
        vals = rng.permutation(n).astype(np.uint32)
        dk = torch.from_numpy(keys.view(np.int32)).cuda()
        dv = torch.from_numpy(vals.view(np.int32)).cuda()
        cnt = torch.tensor([count], dtype=torch.int32, device="cuda")
        s = sb.RadixSorter(ctx, n)
        s.sort(dk.data_ptr(), dv.data_ptr(), cnt.data_ptr(), n, b0, b1)
        torch.cuda.synchronize()
        mask = np.uint32(((1 << (b1 - b0)) - 1) << b0) if b1 - b0 < 32 else np.uint32(0xffffffff)
        order = np.argsort(keys[:count] & mask, kind="stable")
        gk, gv = dk.cpu().numpy().view(np.uint32), dv.cpu().numpy().view(np.uint32)
        assert np.array_equal(gk[:count], keys[:count][order]), (n, b0, b1, count, kind)
        assert np.array_equal(gv[:count], vals[:count][order]), (n, b0, b1, count, kind)
        assert np.array_equal(gk[count:], keys[count:]) and np.array_equal(gv[count:], vals[count:]), "pairs beyond the count must stay"
        s.close()
print("ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SB_SORT_IMPL=impl)
    out = subprocess.run([sys.executable, "-c", code, os.path.join(root, "wgpu-3dgs-viewer_b200")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]


def test_errors(sb, ctx):
    pods = sb.pack_gaussians(sb.scenes.synthetic_gaussians(10, 1))
    v = sb.Viewer(ctx, pods, 10)
    import torch
    t = torch.zeros((16, 16, 4), dtype=torch.uint8, device="cuda")
    v.update_camera((0, 0, -5), 0.0, 0.0, 32, 32)
    with pytest.raises(sb.SplatError):  # target size != CameraPod.size
        v.render(t, 16, 16)
    with pytest.raises(sb.SplatError):
        v.update_gaussian_transform(1.0, 0, 4, False, 3.0)  # GaussianShDegree::new(4) is None
    with pytest.raises(sb.SplatError):
        v.update_gaussian_transform(1.0, 0, 3, False, 3.5)  # GaussianMaxStdDev::new(3.5) is None
    v.close()
    ctx.set_model_size_limit(100)
    with pytest.raises(sb.SplatError) as e:  # ModelSizeExceedsDeviceLimit (preprocessor.rs:239-246)
        sb.Viewer(ctx, pods, 10)
    assert e.value.status == 3
    ctx.set_model_size_limit(1 << 40)


def test_full_size_properties(sb, ctx):
    """BASELINE full size (6M, 1080p): size-independent properties instead of an oracle frame."""
    torch = _torch()
    n, w, h = 6_000_000, 1920, 1080
    g = sb.scenes.synthetic_gaussians(n, sb.scenes.BASE_SEED + 2)
    pods = sb.pack_gaussians(g)
    v = sb.Viewer(ctx, pods, n)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v.update_camera(pos, yaw, pitch, w, h)
    target = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(target, w, h)
    draw, disp = v.read_indirect_args()
    V = int(draw[1])
    assert draw[0] == 6 and disp[0] == (V + 3839) // 3840 and 0 < V <= n
    idx = v.read_indices(V)
    keys = v.read_depth_keys(sb.padded_key_count(n))
    ku = keys[:V].view(np.uint32)
    assert np.all(ku[1:] >= ku[:-1]), "keys not ascending"
    assert np.all(keys[V:int(disp[0]) * 3840] == 2.0)
    assert len(np.unique(idx)) == V and idx.max() < n, "indices are not a set"
    same = ku[1:] == ku[:-1]
    assert np.all(idx[1:][same] > idx[:-1][same]), "ties not in ascending index order"
    # idempotence: a second frame reproduces every artefact bit for bit
    first = target.clone()
    v.render(target, w, h)
    torch.cuda.synchronize()
    assert torch.equal(first, target)
    assert np.array_equal(idx, v.read_indices(V))
    assert not v.read_frame_stats()["overflowed"]
    assert int(first[..., 3].min()) == 255 and int(first[..., :3].max()) > 0
    v.close()


@pytest.mark.parametrize("camera", ["outside", "inside"])
def test_config2_1M_1080p_against_oracle(sb, ob, ctx, camera):
    """BASELINE.json configs[1] at full size — 1 M Gaussians, single/single pods, 1920x1080, splat mode — against the
    oracle: visible set, counts, keys and order bit-exact, framebuffer bit-exact with the reproducible exp and within
    2/255 with the MUFU path (deep per-pixel stacks: ~100 blended layers per pixel with the outside camera)."""
    n, w, h = 1_000_000, 1920, 1080
    g, pods = make_scene(sb, ob, n, sb.scenes.BASE_SEED + 1)
    cam = sb.scenes.CAMERA_OUTSIDE if camera == "outside" else sb.scenes.CAMERA_INSIDE
    r = run_frame(sb, ob, ctx, pods, n, cam, w, h, strict=True)
    check_artifacts(ob, r, n)
    assert not r["stats"]["overflowed"]
    assert img_diff(r) == 0
    fast = run_frame(sb, ob, ctx, pods, n, cam, w, h, strict=False)
    assert img_diff(fast) <= 2


def test_raster_counters_and_stage_timing(sb, ob, ctx):
    """The instrumented rasterizer counts exactly the fragments the oracle blends; stage timers work."""
    torch = _torch()
    n, w, h = 15000, 640, 360
    g, pods = make_scene(sb, ob, n, 66)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n)
    assert v.raster_path() == "tma_gather4"
    v.update_camera(pos, yaw, pitch, w, h)
    v.set_stage_timing(True)
    v.set_raster_counting(True)
    v.set_exact_cutoff(False)  # count every fragment the reference's fragment stage blends, identity blends included
    t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(t, w, h)
    times = v.read_stage_times()
    assert set(times) == set(sb.Viewer.STAGES) and all(ms >= 0 for ms in times.values()) and sum(times.values()) > 0
    c = v.read_raster_counters()
    om = ob.OracleModel(pods, n)
    _, ostats = ob.render(om, ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod())
    assert c["alive"] == ostats["alive_pixels"], "blended fragment count differs from the oracle"
    assert c["evaluated"] >= c["alive"]
    # counting build produces the same image as the normal build
    v.set_raster_counting(False)
    t2 = torch.zeros_like(t)
    v.render(t2, w, h)
    torch.cuda.synchronize()
    assert torch.equal(t, t2)
    v.close()


@pytest.mark.parametrize("camera", ["outside", "inside"])
@pytest.mark.parametrize("strict", [True, False])
def test_exact_alpha_cutoff_is_bit_identical(sb, ob, ctx, camera, strict):
    """The alpha cut-off (splat mode, unorm8) only drops blends that are the identity after re-quantisation: frames with
    and without it are equal bit for bit — with the strict and with the fast exp — while duplicates and evaluated
    fragments shrink; opacities around the threshold (bytes 0..3) and opaque splats are all present."""
    torch = _torch()
    n, w, h = 40000, 800, 450
    g = sb.scenes.synthetic_gaussians(n, 67)
    g["color"][: n // 8, 3] = np.arange(n // 8) % 6          # alpha bytes 0..5: at and around the cut
    g["color"][n // 8: n // 4, 3] = 255
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE if camera == "outside" else sb.scenes.CAMERA_INSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    v.set_strict_exp(strict)
    v.set_raster_counting(True)
    frames, stats, counters = [], [], []
    for cut in (True, False):
        v.set_exact_cutoff(cut)
        t = torch.full((h, w, 4), 3, dtype=torch.uint8, device="cuda")
        v.render(t, w, h)
        torch.cuda.synchronize()
        frames.append(t)
        stats.append(v.read_frame_stats())
        counters.append(v.read_raster_counters())
    assert torch.equal(frames[0], frames[1])
    assert stats[0]["visible"] == stats[1]["visible"]            # the visible set / sort are untouched
    assert stats[0]["duplicates"] < stats[1]["duplicates"]
    assert counters[0]["evaluated"] < counters[1]["evaluated"] and counters[0]["alive"] <= counters[1]["alive"]
    if strict:
        oimg, ost = ob.render(ob.OracleModel(pods, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
        assert np.array_equal(frames[0].cpu().numpy(), oimg)
        assert counters[1]["alive"] == ost["alive_pixels"]
    # float targets never cut: the flag changes nothing there
    vf = sb.Viewer(ctx, pods, n, target_format=sb.TARGET_RGBA32F)
    vf.update_camera(pos, yaw, pitch, w, h)
    dups = []
    for cut in (True, False):
        vf.set_exact_cutoff(cut)
        tf = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        vf.render(tf, w, h)
        torch.cuda.synchronize()
        dups.append(vf.read_frame_stats()["duplicates"])
    assert dups[0] == dups[1] == stats[1]["duplicates"]
    v.close(); vf.close()


@pytest.mark.parametrize("strict", [True, False])
def test_exact_alpha_cutoff_adversarial(sb, ob, ctx, strict):
    """The cut-off's rounding argument at its edge: screen-filling splats whose alpha sweeps continuously through the
    threshold, pure black over pure white and pure white over black (|c - d| = 255, the worst case of the bound), opacity
    bytes 1..4 as well as opaque ones.  Cut-off and no-discard build on/off give the same frame bit for bit; the strict frame
    equals the oracle's."""
    torch = _torch()
    w, h = 320, 180
    rng = np.random.default_rng(11)
    n = 400
    g = np.zeros(n, dtype=sb.GAUSSIAN_DTYPE)
    g["rot"][:, 3] = 1.0
    # all in front of the "outside" camera, near the view axis, big enough to cover the frame with slowly varying r^2
    g["pos"][:, 0] = rng.uniform(-4, 4, n)
    g["pos"][:, 1] = rng.uniform(-3, 3, n)
    g["pos"][:, 2] = np.linspace(8, -8, n)            # far to near in index order too
    g["scale"][:] = rng.uniform(1.0, 6.0, (n, 1))
    g["scale"][:, 1] *= rng.uniform(0.3, 1.0, n)       # some anisotropy
    white = (np.arange(n) // 7) % 2 == 0
    g["color"][white, :3] = 255
    g["color"][~white, :3] = 0
    a = rng.choice([1, 2, 3, 4, 8, 40, 255], n, p=[0.2, 0.2, 0.15, 0.1, 0.1, 0.1, 0.15])
    g["color"][:, 3] = a
    g["color"][::50, 3] = 255                         # opaque layers that flip the destination to the other extreme
    pods = sb.pack_gaussians(g)
    pos, yaw, pitch = sb.scenes.CAMERA_OUTSIDE
    v = sb.Viewer(ctx, pods, n)
    v.update_camera(pos, yaw, pitch, w, h)
    v.update_gaussian_transform(1.0, sb.MODE_SPLAT, 0, False, 3.0)   # sh_deg 0: colours exactly 0 / 255
    v.set_strict_exp(strict)
    frames = []
    for cut in (True, False):
        v.set_exact_cutoff(cut)
        t = torch.full((h, w, 4), 9, dtype=torch.uint8, device="cuda")
        v.render(t, w, h)
        torch.cuda.synchronize()
        frames.append(t.cpu().numpy())
    assert np.array_equal(frames[0], frames[1])
    assert frames[0][..., :3].min() < 30 and frames[0][..., :3].max() > 200   # both extremes are on screen
    if strict:
        oimg, _ = ob.render(ob.OracleModel(pods, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(1.0, 0, 0, False, 3.0),
                            strict_exp=True)
        assert np.array_equal(frames[0], oimg)
    v.close()


def test_bulk_raster_path_parity(sb, ob):
    """SB_RASTER_PATH=bulk (gathered copy + 1-D TMA bulk copies) stays bit-identical to the default
    TMA gather4 path; run in a subprocess because the path is chosen at viewer creation."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import splat_b200 as sb
from oracle import binding as ob
n, w, h = 12000, 512, 288
g = sb.scenes.synthetic_gaussians(n, 91)
pods = sb.pack_gaussians(g)
ctx = sb.Context(0)
v = sb.Viewer(ctx, pods, n)
assert v.raster_path() == "bulk", v.raster_path()
v.set_strict_exp(True)
pos, yaw, pitch = sb.scenes.CAMERA_INSIDE
v.update_camera(pos, yaw, pitch, w, h)
t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
v.render(t, w, h)
torch.cuda.synchronize()
oimg, _ = ob.render(ob.OracleModel(pods, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
d = np.abs(t.cpu().numpy().astype(np.int32) - oimg.astype(np.int32)).max()
print("maxdiff", d)
assert d == 0
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SB_RASTER_PATH="bulk")
    out = subprocess.run([sys.executable, "-c", code, os.path.join(root, "wgpu-3dgs-viewer_b200"), root], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


def test_bbox_only_cull_is_bit_identical(sb, ob):
    """SB_RASTER_CULL=bbox switches the warp-level cull back to the alive-region bbox alone; the separating-axis cull
    of the default path may only drop (patch, splat) pairs with no alive fragment, so the frames are identical."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import splat_b200 as sb
from oracle import binding as ob
n, w, h = 20000, 640, 360
g = sb.scenes.synthetic_gaussians(n, 77, log_scale=(-5.0, -2.0))
pods = sb.pack_gaussians(g)
ctx = sb.Context(0)
for cam in (sb.scenes.CAMERA_INSIDE, sb.scenes.CAMERA_OUTSIDE):
    v = sb.Viewer(ctx, pods, n)
    v.set_strict_exp(True)
    v.set_raster_counting(True)
    v.set_exact_cutoff(False)
    pos, yaw, pitch = cam
    v.update_camera(pos, yaw, pitch, w, h)
    t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    v.render(t, w, h)
    c = v.read_raster_counters()
    oimg, ost = ob.render(ob.OracleModel(pods, n), ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod(), strict_exp=True)
    d = np.abs(t.cpu().numpy().astype(np.int32) - oimg.astype(np.int32)).max()
    print("maxdiff", d, c, ost)
    assert d == 0 and c["alive"] == ost["alive_pixels"]
    v.close()
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SB_RASTER_CULL="bbox")
    out = subprocess.run([sys.executable, "-c", code, os.path.join(root, "wgpu-3dgs-viewer_b200"), root], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


def test_render_batch_equals_single_frames(sb, ob, ctx):
    """sb_viewer_render_batch (views pipelined two deep on internal streams) is bit-identical to
    update_camera + render per view, for device targets and for host frames; artefacts stay exact."""
    torch = _torch()
    n, w, h = 30000, 640, 360
    g, pods = make_scene(sb, ob, n, 44)
    v = sb.Viewer(ctx, pods, n)
    rng = np.random.default_rng(8)
    sel = rng.integers(0, 2**32, size=(n + 31) // 32, dtype=np.uint64).astype(np.uint32)
    v.enable_selection(True)
    v.set_selection(sel)
    v.update_model_transform((0.5, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), (1.1, 1.0, 0.9))
    cams = [sb.camera_pod(*sb.scenes.orbit_camera(k, 7), w, h) for k in range(7)]
    singles = []
    for c in cams:
        t = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        v.update_camera_with_pod(c)
        v.render(t, w, h)
        singles.append(t)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    outs = [torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda") for _ in cams]
    v.render_batch(cams, targets=outs, width=w, height=h, stream=stream)
    stream.synchronize()
    for a, b in zip(singles, outs):
        assert torch.equal(a, b)
    hosts = [torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory() for _ in cams]
    v.render_batch(cams, host_ptrs=[t.data_ptr() for t in hosts], stream=stream)
    stream.synchronize()
    for a, b in zip(singles, hosts):
        assert torch.equal(a.cpu(), b)
    # the primary viewer's artefacts belong to the last even-indexed view (slot 0) and stay oracle-exact
    om = ob.OracleModel(pods, n, model_transform=ob.model_transform_pod((0.5, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), (1.1, 1.0, 0.9)),
                        selection=sel, invert_selection=1)
    pos, yaw, pitch = sb.scenes.orbit_camera(6, 7)
    pre = ob.preprocess(om, ob.camera_pod(pos, yaw, pitch, w, h), ob.gaussian_transform_pod())
    draw, _ = v.read_indirect_args(stream)
    assert int(draw[1]) == pre["count"]
    _, oi = ob.radix_sort(pre["keys"][: pre["count"]].view(np.uint32), pre["indices"][: pre["count"]])
    assert np.array_equal(v.read_indices(pre["count"], stream), oi)
    v.close()
