"""CPU: the C-ABI library loads, exports every symbol include/splat_b200.h declares, and its
host-only helpers agree with the oracle.  No compute entry point is called (no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest


def header_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "splat_b200.h")).read()
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(sb):
    lib = sb.load()
    declared = header_symbols()
    assert len(declared) >= 60
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/splat_b200.h but not exported"
    assert sorted(sb.EXPORTED_SYMBOLS) == declared, "api.py binding list out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", sb.lib_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared) <= exported
    leaked = [s for s in exported if not s.startswith("sb_")]
    assert not leaked, f"non-ABI symbols exported: {leaked[:5]}"


def test_library_is_sm100a_cuda_and_has_no_oracle_dependency(sb):
    out = subprocess.run(["cuobjdump", "-lelf", sb.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    needed = subprocess.run(["readelf", "-d", sb.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in needed


def test_pod_layout_sizes(sb, ob):
    # reference: buffer sizes == size_of::<Pod>() (tests/buffer/camera.rs, indirect_args.rs)
    assert ctypes.sizeof(sb.CameraPod) == 144
    assert ctypes.sizeof(sb.ModelTransformPod) == 48
    assert ctypes.sizeof(sb.GaussianTransformPod) == 8
    expect = {(0, 0): 224, (0, 1): 208, (0, 2): 224, (1, 0): 144, (1, 1): 128, (1, 2): 144, (2, 0): 96, (2, 1): 80,
              (2, 2): 96, (3, 0): 48, (3, 1): 32, (3, 2): 48}
    for (sh, cov), stride in expect.items():
        assert sb.pod_stride(sh, cov) == stride == ob.pod_stride(sh, cov)
    assert sb.pod_stride(7, 0) == 0
    # wgpu_sort::keys_buffer_size_bytes (radix_sorter.rs:922-939)
    for n in (0, 1, 3839, 3840, 3841, 1_000_000, 6_000_000):
        assert sb.padded_key_count(n) == (n + 3839) // 3840 * 3840
        assert sb.keys_buffer_size_bytes(n) == sb.padded_key_count(n) * 16


def test_host_packer_matches_oracle(sb, ob):
    g = sb.scenes.synthetic_gaussians(2000, 4)
    g["sh"][5] = 0.0  # degenerate norm8 range
    g["sh"][6] = 0.25
    for sh in range(4):
        for cov in range(3):
            a = sb.pack_gaussians(g, sh, cov)
            b = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE), sh, cov)
            assert np.array_equal(a, b), (sh, cov)


def test_host_packer_half_conversion_special_values(sb, ob):
    """The packer converts to half eight at a time in hardware where the host has F16C; the oracle converts in software.  Every
    class of value must come out the same: ties to even, the overflow threshold (65520 -> inf), denormal halves and the underflow
    threshold (2^-25 ties to zero), signed zeros, infinities and NaNs (payload dropped), next to a dense random sweep of exponents."""
    rng = np.random.default_rng(21)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 65504.0, 65519.99, 65520.0, 65536.0, 1e9, -1e9, np.inf, -np.inf, np.nan,
                         2.0 ** -24, 2.0 ** -25, np.nextafter(np.float32(2.0 ** -25), np.float32(1.0)), 2.0 ** -26, 6.0e-8, 6.1e-5, 6.103515625e-05,
                         1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 1.0 + 2.0 ** -11 + 2.0 ** -20, 1.0 + 2.0 ** -11 - 2.0 ** -20, 2047.5, 2048.5,
                         0.333333343, -0.1, 3.14159274, 1e-40, -1e-40], dtype=np.float32)
    nanp = np.array([0x7fc12345, 0xffc00001, 0x7f800001], dtype=np.uint32).view(np.float32)  # NaNs with payloads
    n = 4096
    g = sb.scenes.synthetic_gaussians(n, 22)
    sweep = (rng.standard_normal((n, 45)) * np.exp2(rng.integers(-30, 18, (n, 45)))).astype(np.float32)
    g["sh"] = sweep
    flat = np.concatenate([specials, nanp])
    g["sh"].reshape(-1)[: len(flat)] = flat
    g["sh"][100, :8] = specials[:8]          # specials inside a full vector of eight as well as in a tail
    g["sh"][101, 37:45] = specials[8:16]
    g["scale"][:16] = np.exp2(rng.integers(-14, 8, (16, 3))).astype(np.float32)   # covariances across the half range (cov half)
    for sh, cov in ((1, 0), (1, 1), (0, 1), (2, 1)):
        a = sb.pack_gaussians(g, sh, cov)
        b = ob.pack_gaussians(g.view(ob.GAUSSIAN_DTYPE), sh, cov)
        assert np.array_equal(a, b), (sh, cov)


def test_camera_and_transform_pods_match_oracle(sb, ob):
    rng = np.random.default_rng(0)
    for _ in range(50):
        pos = rng.uniform(-20, 20, 3)
        yaw, pitch = rng.uniform(0, 6.28), rng.uniform(-1.5, 1.5)
        w, h = int(rng.integers(16, 4000)), int(rng.integers(16, 3000))
        assert bytes(sb.camera_pod(pos, yaw, pitch, w, h)) == bytes(ob.camera_pod(pos, yaw, pitch, w, h))
    # CameraPod::new semantics (tests/buffer/camera.rs:95-125): yaw = 0 looks down +Z, depth 0..1
    pod = sb.camera_pod((0, 0, 0), 0.0, 0.0, 100, 100)
    view = np.array(pod.view).reshape(4, 4).T
    assert np.allclose(view @ np.array([0, 0, 1, 1.0]), [0, 0, -1, 1])
    proj = np.array(pod.proj).reshape(4, 4).T
    near = proj @ np.array([0, 0, -0.1, 1.0])
    far = proj @ np.array([0, 0, -1e4, 1.0])
    assert abs(near[2] / near[3]) < 1e-6 and abs(far[2] / far[3] - 1) < 1e-4
    assert bytes(sb.model_transform_pod((1, 2, 3), (0, 0, 0, 1), (2, 2, 2))) == bytes(ob.model_transform_pod((1, 2, 3), (0, 0, 0, 1), (2, 2, 2)))
    for sd in (0.0, 1.0, 2.5, 3.0):
        assert bytes(sb.gaussian_transform_pod(1.5, 1, 2, True, sd)) == bytes(ob.gaussian_transform_pod(1.5, 1, 2, True, sd))


def test_ply_reader_matches_oracle(sb, ob, tmp_path):
    here = os.path.dirname(os.path.abspath(__file__))
    props = np.load(os.path.join(here, "golden", "model_ply_props.npy"))
    names = (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)]
             + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(props)
    hdr += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    path = tmp_path / "model.ply"
    path.write_bytes(hdr.encode() + props.astype("<f4").tobytes())
    g = sb.read_ply(str(path))
    og = ob.gaussians_from_ply_props(props)
    assert g.tobytes() == og.tobytes()
    assert len(g) == 9 and g["color"][0].tolist() == [255, 0, 0, 255]


def test_ply_reader_block_boundaries(sb, ob, tmp_path):
    """sb_read_ply converts the rows in 65536-row blocks on all host cores: a file that spans a full block plus a ragged one (and
    one of exactly one block) comes out byte for byte as the oracle's row-by-row conversion."""
    names = (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)]
             + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])
    rng = np.random.default_rng(33)
    for n in (65536, 65536 + 4465):
        props = rng.standard_normal((n, len(names))).astype("<f4")
        props[:, names.index("opacity")] *= 4.0
        hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
        hdr += "".join(f"property float {x}\n" for x in names) + "end_header\n"
        path = tmp_path / f"blocks_{n}.ply"
        path.write_bytes(hdr.encode() + props.tobytes())
        g = sb.read_ply(str(path))
        og = ob.gaussians_from_ply_props(props)
        assert len(g) == n and g.tobytes() == og.tobytes()


def test_no_context_without_gpu(sb):
    import torch
    if torch.cuda.is_available():
        return
    try:
        sb.Context(0)
    except sb.SplatError as e:
        assert e.status == 2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("sb_ctx_create must fail loudly without a CUDA device")


def _write_ply(path, names, rows, vertex_count=None):
    hdr = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % (len(rows) if vertex_count is None else vertex_count)
    hdr += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    path.write_bytes(hdr.encode() + np.asarray(rows, dtype="<f4").tobytes())


def test_ply_reader_rejects_incomplete_or_lying_files(sb, tmp_path):
    """sb_read_ply must fail with SB_ERR_IO — never read out of bounds — when a required property is missing, the f_rest set is
    partial in a way that is not a whole SH degree, or the header's vertex count exceeds the file."""
    base = ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((5, len(base))).astype(np.float32)
    p = tmp_path / "ok.ply"
    _write_ply(p, base, rows)
    g = sb.read_ply(str(p))
    assert len(g) == 5 and np.all(g["sh"] == 0)
    for missing in ("f_dc_1", "f_dc_2", "scale_2", "rot_3", "rot_1", "opacity", "z"):
        names = [n for n in base if n != missing]
        _write_ply(p, names, rows[:, : len(names)])
        with pytest.raises(sb.SplatError) as e:
            sb.read_ply(str(p))
        assert e.value.status == 6, missing
    # vertex count far beyond the data (and absurdly large): rejected before any allocation
    for count in (6, 10**15):
        _write_ply(p, base, rows, vertex_count=count)
        with pytest.raises(sb.SplatError):
            sb.read_ply(str(p))
    # f_rest sets: 9 properties = SH degree 1 (3 per channel, channel-major) is read as such; 10 or a holey set is rejected
    names9 = base + [f"f_rest_{i}" for i in range(9)]
    rows9 = rng.standard_normal((4, len(names9))).astype(np.float32)
    _write_ply(p, names9, rows9)
    g = sb.read_ply(str(p))
    rest = rows9[:, len(base):]
    for k in range(3):
        for c in range(3):
            assert np.array_equal(g["sh"][:, k * 3 + c], rest[:, c * 3 + k])
    assert np.all(g["sh"][:, 9:] == 0)
    for bad in (base + [f"f_rest_{i}" for i in range(10)], base + ["f_rest_0", "f_rest_2", "f_rest_3"]):
        _write_ply(p, bad, rng.standard_normal((2, len(bad))).astype(np.float32))
        with pytest.raises(sb.SplatError):
            sb.read_ply(str(p))


def test_rust_binding_crate_binds_and_wraps_every_export():
    """The Rust shim (source only: no cargo here) must stay in step with the header: the generated `extern "C"` block is up
    to date, declares every `SB_API` export, and the safe layer calls each of them (36 of 79 were unbound)."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, "rust/src/ffi_gen.rs is stale: run scripts/gen_rust_ffi.py\n" + r.stdout + r.stderr
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(root, "include", "splat_b200.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(sb_\w+)\s*\(", hdr)))
    gen = open(os.path.join(root, "wgpu-3dgs-viewer_b200", "rust", "src", "ffi_gen.rs")).read()
    lib = open(os.path.join(root, "wgpu-3dgs-viewer_b200", "rust", "src", "lib.rs")).read()
    assert len(names) >= 96
    assert [n for n in names if f"pub fn {n}(" not in gen] == []
    assert [n for n in names if f"ffi::{n}(" not in lib] == [], "exports without a safe wrapper in rust/src/lib.rs"
    assert 'include!("ffi_gen.rs")' in lib


@pytest.mark.parametrize("version,sh_degree,frac_bits", [(2, 3, 12), (3, 3, 12), (2, 0, 10), (3, 2, 14), (2, 1, 12)])
def test_read_spz_against_the_numpy_reader(sb, ob, tmp_path, version, sh_degree, frac_bits):
    """GaussiansSource::Spz (examples/simple.rs:157-160): the C++ reader and the independent numpy reader decode the same
    Gaussians from fixtures written by the numpy writer — v2 (3-byte rotations) and v3 (smallest-three), every SH degree,
    negative coordinates, both signs of every quaternion component; truncated and foreign files are rejected."""
    from oracle import spz_np
    rng = np.random.default_rng(100 + version * 10 + sh_degree)
    n = 777
    pos = rng.uniform(-40, 40, (n, 3))
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[:4] = np.eye(4)          # exact axes: a zero smallest-three residue, and w = 0 for the v2 reconstruction
    q[4:8] = -np.eye(4)
    path = str(tmp_path / "scene.spz")
    spz_np.write_spz(path, pos, rng.integers(0, 256, n), rng.integers(0, 256, (n, 3)), rng.integers(0, 256, (n, 3)), q,
                     rng.integers(0, 256, (n, 15, 3)), version=version, frac_bits=frac_bits, sh_degree=sh_degree)
    got = sb.read_spz(path)
    exp = spz_np.read_spz(path, sb.GAUSSIAN_DTYPE)
    assert len(got) == n
    for f in ("pos", "color", "sh", "rot"):
        assert np.array_equal(got[f], exp[f]), f
    assert np.allclose(got["scale"], exp["scale"], rtol=2e-7, atol=0)  # exp(): libm vs numpy, <= 1 ulp
    assert np.abs(got["pos"] - pos).max() <= 0.5 / (1 << frac_bits) + 1e-6
    # same rotation up to sign and the quantisation of the container (8-bit components in v2, 9-bit magnitudes in v3)
    assert np.abs(np.abs(np.sum(got["rot"] * q, axis=1)) - 1).max() < (6e-3 if version == 2 else 2e-3)
    dim = (0, 3, 8, 15)[sh_degree]
    assert np.all(got["sh"].reshape(n, 15, 3)[:, dim:] == 0)
    # a pod packed from the decoded Gaussians renders through the same path as any other source
    assert sb.pack_gaussians(got).shape[0] == n * sb.pod_stride(0, 0) or sb.pack_gaussians(got).size == n * sb.pod_stride(0, 0)
    # rejects: not gzip / wrong magic / truncated body
    bad = tmp_path / "bad.spz"
    bad.write_bytes(b"not a gzip stream at all")
    with pytest.raises(sb.SplatError):
        sb.read_spz(str(bad))
    import gzip
    raw = gzip.open(path, "rb").read()
    with gzip.open(str(bad), "wb") as f:
        f.write(raw[: len(raw) // 2])
    with pytest.raises(sb.SplatError):
        sb.read_spz(str(bad))
    with gzip.open(str(bad), "wb") as f:
        f.write(b"XXXX" + raw[4:])
    with pytest.raises(sb.SplatError):
        sb.read_spz(str(bad))
