"""Niantic .spz (versions 2 and 3) in numpy — TEST INFRASTRUCTURE ONLY: an independent reader (checks the product's
sb_read_spz field by field) and a writer (makes the fixtures).  Layout as published with the format (RECALLED; the decoder the
reference uses lives in the un-vendored wgpu-3dgs-core, examples/simple.rs:157-160 only names GaussiansSource::Spz):
gzip( header{u32 'NGSP', u32 version, u32 numPoints, u8 shDegree, u8 fractionalBits, u8 flags, u8 reserved}
      positions[n][3] as 24-bit signed fixed point | alphas[n] u8 | colors[n][3] u8 | scales[n][3] u8
      rotations[n][3] u8 (v2) or [n] u32 smallest-three (v3) | sh[n][shDim][3] u8 ),  shDim = 0, 3, 8, 15."""
from __future__ import annotations

import gzip
import struct

import numpy as np

f32 = np.float32
SH_DIM = (0, 3, 8, 15)
MAGIC = 0x5053474E


def write_spz(path, pos, alpha_u8, color_u8, scale_u8, rot_xyzw, sh_u8, version=2, frac_bits=12, sh_degree=3):
    """Quantised inputs in, file out.  pos: float (rounded to the fixed-point grid), rot_xyzw: unit quaternions."""
    n = len(pos)
    fixed = np.rint(np.asarray(pos, dtype=np.float64) * (1 << frac_bits)).astype(np.int64)
    assert np.all(np.abs(fixed) < (1 << 23))
    u = (fixed & 0xFFFFFF).astype(np.uint32)
    pb = np.stack([(u & 255), (u >> 8) & 255, (u >> 16) & 255], axis=-1).astype(np.uint8).reshape(n, 9)
    q = np.asarray(rot_xyzw, dtype=np.float64)
    if version == 2:
        q = q * np.where(q[:, 3:4] < 0, -1.0, 1.0)  # w is reconstructed as the positive root
        rb = np.clip(np.rint((q[:, :3] + 1.0) * 127.5), 0, 255).astype(np.uint8)
    else:
        largest = np.argmax(np.abs(q), axis=1)
        q = q * np.where(np.take_along_axis(q, largest[:, None], 1) < 0, -1.0, 1.0)
        comp = largest.astype(np.uint64)
        for k in range(4):  # pack in ascending k, so the reader peels them off in descending k
            use = largest != k
            mag = np.clip(np.rint(np.abs(q[:, k]) / 0.70710678 * 511.0), 0, 511).astype(np.uint64)
            neg = (q[:, k] < 0).astype(np.uint64)
            comp = np.where(use, (comp << np.uint64(10)) | (neg << np.uint64(9)) | mag, comp)
        rb = comp.astype(np.uint32).view(np.uint8).reshape(n, 4)
    dim = SH_DIM[sh_degree]
    body = (pb.tobytes() + np.asarray(alpha_u8, np.uint8).tobytes() + np.asarray(color_u8, np.uint8).reshape(n, 3).tobytes()
            + np.asarray(scale_u8, np.uint8).reshape(n, 3).tobytes() + rb.tobytes()
            + np.asarray(sh_u8, np.uint8).reshape(n, 15, 3)[:, :dim].tobytes())
    hdr = struct.pack("<IIIBBBB", MAGIC, version, n, sh_degree, frac_bits, 0, 0)
    with gzip.open(path, "wb") as f:
        f.write(hdr + body)


def read_spz(path, gaussian_dtype):
    """-> structured array (pos, color, sh, scale, rot) with the same mapping as the PLY path (colour from the DC term)."""
    raw = gzip.open(path, "rb").read()
    magic, version, n, sh_degree, frac_bits, _flags, _res = struct.unpack("<IIIBBBB", raw[:16])
    assert magic == MAGIC and version in (2, 3)
    dim = SH_DIM[sh_degree]
    a = np.frombuffer(raw, dtype=np.uint8, offset=16)
    o = 0

    def take(count):
        nonlocal o
        v = a[o:o + count]
        o += count
        return v

    pb = take(n * 9).reshape(n, 3, 3).astype(np.int32)
    alpha = take(n)
    col = take(n * 3).reshape(n, 3)
    scl = take(n * 3).reshape(n, 3)
    rotb = take(n * (4 if version == 3 else 3))
    shb = take(n * dim * 3).reshape(n, dim, 3)
    g = np.zeros(n, dtype=gaussian_dtype)
    v = pb[..., 0] | (pb[..., 1] << 8) | (pb[..., 2] << 16)
    v = np.where(v & 0x800000, v - (1 << 24), v)
    g["pos"] = v.astype(f32) * (f32(1.0) / f32(1 << frac_bits))
    dc = (col.astype(f32) / f32(255.0) - f32(0.5)) / f32(0.15)
    c = (f32(0.5) + f32(0.2820948) * dc) * f32(255.0)
    g["color"][:, :3] = np.clip(c, 0, 255).astype(np.uint8)  # truncation, as Rust `as u8`
    g["color"][:, 3] = alpha
    g["scale"] = np.exp((scl.astype(f32) / f32(16.0) - f32(10.0)).astype(f32)).astype(f32)
    q = np.zeros((n, 4), dtype=f32)
    if version == 2:
        q[:, :3] = rotb.reshape(n, 3).astype(f32) / f32(127.5) - f32(1.0)
        s = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]
        q[:, 3] = np.sqrt(np.maximum(f32(0.0), f32(1.0) - s))
    else:
        comp = rotb.view(np.uint32).copy()
        largest = (comp >> 30).astype(np.int64)
        s = np.zeros(n, dtype=f32)
        for k in (3, 2, 1, 0):
            use = largest != k
            mag = (comp & 0x1FF).astype(f32)
            neg = ((comp >> 9) & 1).astype(bool)
            val = f32(0.70710678) * mag / f32(511.0)
            val = np.where(neg, -val, val).astype(f32)
            q[:, k] = np.where(use, val, q[:, k])
            s = np.where(use, s + val * val, s).astype(f32)
            comp = np.where(use, comp >> 10, comp)
        q[np.arange(n), largest] = np.sqrt(np.maximum(f32(0.0), f32(1.0) - s))
    ln = np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3])
    inv = np.where(ln > 0, f32(1.0) / ln, f32(0.0)).astype(f32)
    g["rot"] = q * inv[:, None]
    g["rot"][ln <= 0, 3] = 1.0
    sh = np.zeros((n, 15, 3), dtype=f32)
    sh[:, :dim] = (shb.astype(f32) - f32(128.0)) / f32(128.0)
    g["sh"] = sh.reshape(n, 45)
    return g
