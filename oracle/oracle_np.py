"""numpy twin of the C oracle's preprocess + vertex stage (TEST INFRASTRUCTURE ONLY).

A second, independently written statement of the same arithmetic contract (every f32 op rounded
on its own, order as in the WGSL text) used to cross-check oracle/splat_oracle.c bit for bit on
small and medium scenes.  Vectorised over Gaussians; numpy float32 ufuncs are individually
rounded IEEE operations (no matmul/BLAS, no FMA).

Restates: preprocess.wesl:60-106, utils.wesl:18-79, camera.wesl:13-15 (reference src/shader/).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _mat3_from_quat(q):
    x, y, z, w = (f32(v) for v in q)
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    one = f32(1.0)
    return np.array([[one - (yy + zz), xy + wz, xz - wy],
                     [xy - wz, one - (xx + zz), yz + wx],
                     [xz + wy, yz - wx, one - (xx + yy)]], dtype=f32)  # [col][row]


def _mat4_mul(a, b):  # column-major [col][row]
    r = np.zeros((4, 4), dtype=f32)
    for j in range(4):
        for i in range(4):
            r[j, i] = ((a[0, i] * b[j, 0] + a[1, i] * b[j, 1]) + a[2, i] * b[j, 2]) + a[3, i] * b[j, 3]
    return r


class Uniforms:
    def __init__(self, cam, mt, gt):
        self.view = np.array(cam.view, dtype=f32).reshape(4, 4)
        self.proj = np.array(cam.proj, dtype=f32).reshape(4, 4)
        self.size = np.array(cam.size, dtype=f32)
        rot = np.array(mt.rot, dtype=f32)
        scale = np.array(mt.scale, dtype=f32)
        pos = np.array(mt.pos, dtype=f32)
        r = _mat3_from_quat(rot)
        self.sr = np.zeros((3, 3), dtype=f32)
        for c in range(3):
            for row in range(3):
                self.sr[c, row] = r[c, row] * scale[c]
        self.model = np.zeros((4, 4), dtype=f32)
        self.model[:3, :3] = self.sr
        self.model[3, :3] = pos
        self.model[3, 3] = 1.0
        self.pv = _mat4_mul(self.proj, self.view)
        self.vm = _mat4_mul(self.view, self.model)
        self.w = self.view[:3, :3].copy()
        self.focal = np.array([self.proj[0, 0] * self.size[0] * f32(0.5), self.proj[1, 1] * self.size[1] * f32(0.5)], dtype=f32)
        self.std_dev = f32(gt.max_std_dev) / f32(255.0) * f32(3.0)
        self.gsize = f32(gt.size)


def _cull(x, y, z):
    with np.errstate(invalid="ignore"):
        return ~((x >= -1) & (y >= -1) & (z >= 0) & (x <= 1) & (y <= 1) & (z <= 1))


def cov2d_axes(u: Uniforms, pos, cov3d, std_dev):
    """pos: (n,3) f32, cov3d: (n,6) f32 -> axes (n,4) f32."""
    px, py, pz = pos[:, 0], pos[:, 1], pos[:, 2]
    vm = u.vm
    t = [((vm[0, i] * px + vm[1, i] * py) + vm[2, i] * pz) + vm[3, i] for i in range(3)]
    with np.errstate(all="ignore"):
        tz2 = t[2] * t[2]
        j00 = u.focal[0] / t[2]
        j02 = -(u.focal[0] * t[0]) / tz2
        j11 = u.focal[1] / t[2]
        j12 = -(u.focal[1] * t[1]) / tz2
        jw0 = [j00 * u.w[c, 0] + j02 * u.w[c, 2] for c in range(3)]
        jw1 = [j11 * u.w[c, 1] + j12 * u.w[c, 2] for c in range(3)]
        t0 = [(jw0[0] * u.sr[c, 0] + jw0[1] * u.sr[c, 1]) + jw0[2] * u.sr[c, 2] for c in range(3)]
        t1 = [(jw1[0] * u.sr[c, 0] + jw1[1] * u.sr[c, 1]) + jw1[2] * u.sr[c, 2] for c in range(3)]
        v = [[cov3d[:, 0], cov3d[:, 1], cov3d[:, 2]], [cov3d[:, 1], cov3d[:, 3], cov3d[:, 4]], [cov3d[:, 2], cov3d[:, 4], cov3d[:, 5]]]
        tv0 = [(t0[0] * v[0][c] + t0[1] * v[1][c]) + t0[2] * v[2][c] for c in range(3)]
        tv1 = [(t1[0] * v[0][c] + t1[1] * v[1][c]) + t1[2] * v[2][c] for c in range(3)]
        ca = (tv0[0] * t0[0] + tv0[1] * t0[1]) + tv0[2] * t0[2]
        cb = (tv1[0] * t0[0] + tv1[1] * t0[1]) + tv1[2] * t0[2]
        cc = (tv1[0] * t1[0] + tv1[1] * t1[1]) + tv1[2] * t1[2]
        mid = f32(0.5) * (ca + cc)
        hx = f32(0.5) * (ca - cc)
        radius = np.sqrt(hx * hx + cb * cb)
        lam1 = mid + radius
        lam2 = mid - radius
        dx, dy = cb, lam1 - ca
        l = np.sqrt(dx * dx + dy * dy)
        zero = (dx == 0) & (dy == 0)
        ddx = np.where(zero, f32(0.0), dx / l)
        ddy = np.where(zero, f32(1.0), dy / l)
        major = np.minimum(std_dev * np.sqrt(lam1), f32(1024.0))
        minor = np.minimum(std_dev * np.sqrt(lam2), f32(1024.0))
        # fminf semantics: NaN operand -> the other operand
        major = np.where(np.isnan(std_dev * np.sqrt(lam1)), f32(1024.0), major)
        minor = np.where(np.isnan(std_dev * np.sqrt(lam2)), f32(1024.0), minor)
        axes = np.stack([major * ddx, major * ddy, minor * ddy, minor * -ddx], axis=1).astype(f32)
        axes[lam2 < 0] = 0.0
    return axes


def preprocess(pos, cov3d, cam, mt, gt, selection=None, invert=1):
    """Returns (visible bool mask, keys f32 for all Gaussians, ndc)."""
    u = Uniforms(cam, mt, gt)
    pos = np.ascontiguousarray(pos, dtype=f32)
    n = len(pos)
    px, py, pz = pos[:, 0], pos[:, 1], pos[:, 2]
    m = u.model
    world = [((m[0, i] * px + m[1, i] * py) + m[2, i] * pz) + m[3, i] for i in range(3)]
    pv = u.pv
    clip = [((pv[0, i] * world[0] + pv[1, i] * world[1]) + pv[2, i] * world[2]) + pv[3, i] for i in range(4)]
    with np.errstate(all="ignore"):
        nx, ny, nz = clip[0] / clip[3], clip[1] / clip[3], clip[2] / clip[3]
        vis = np.ones(n, dtype=bool)
        if selection is not None:
            idx = np.arange(n, dtype=np.uint32)
            bit = ((selection[idx >> 5] >> (idx & 31)) & 1).astype(bool)
            vis &= ~((invert != 0) == bit)
        culled = _cull(nx, ny, nz)
        axes = cov2d_axes(u, pos, np.ascontiguousarray(cov3d, dtype=f32), u.std_dev * u.gsize)
        mx = axes[:, 0] * u.std_dev / u.size[0]
        my = axes[:, 1] * u.std_dev / u.size[1]
        ndc_major_len = np.sqrt(mx * mx + my * my)
        l = np.sqrt(nx * nx + ny * ny)
        dirx, diry = -nx / l, -ny / l
        mm = np.where(np.isnan(ndc_major_len), l, np.where(np.isnan(l), ndc_major_len, np.minimum(ndc_major_len, l)))
        bx, by = nx + mm * dirx, ny + mm * diry
        culled2 = _cull(bx, by, nz)
        vis &= ~(culled & culled2)
        keys = (f32(1.0) - nz).astype(f32)
    return vis, keys, (nx, ny, nz)
