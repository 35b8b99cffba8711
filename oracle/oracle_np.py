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


# =====================================================================================================================
# Draw stage, second statement.  Two independent restatements of render.wesl + utils.wesl:view_color:
#
#   (1) project_f32 / composite_f32 — the strict-f32 contract (every op individually rounded, fma where DESIGN.md §3 says
#       so), vectorised numpy; must agree with oracle/splat_oracle.c BIT FOR BIT (so_project / so_draw with strict_exp).
#       numpy has no fma: fma32() emulates it exactly (exact f64 product, TwoSum-corrected single rounding).
#   (2) draw_reference_f64 — written from the WGSL text alone, in float64 and in the shader's own formulation: clip-space
#       quad corners, the hardware's affine interpolation of `quad_offset` across the quad (solved per pixel from the
#       2x2 corner matrix), `exp`, ALPHA_BLENDING with RNE re-quantisation of the unorm8 target after every blend and the
#       fixed-point source clamp.  It shares no code, no evaluation order and no pre-inverted axes with the C oracle; the
#       two must agree within the framebuffer tolerance (2/255) and on almost every pixel exactly.
#
# Restates: render.wesl:46-181, utils.wesl:25-135, camera.wesl:1-15 (reference src/shader/).
# =====================================================================================================================

f64 = np.float64

SH_BYTES = {0: 180, 1: 92, 2: 52, 3: 0}       # single, half, norm8, none
COV_BYTES = {0: 24, 1: 12, 2: 28}             # single, half, rot+scale


def pod_stride(sh_fmt: int, cov_fmt: int) -> int:
    return (16 + SH_BYTES[sh_fmt] + COV_BYTES[cov_fmt] + 15) & ~15


def unpack_pods(pods: np.ndarray, n: int, sh_fmt: int = 0, cov_fmt: int = 0):
    """Packed GaussianPod bytes -> (pos (n,3) f32, color (n,4) u8, sh (n,45) f32, cov3d (n,6) f32): the core crate's
    gaussian_unpack_{color,sh,cov3d} as SURVEY Appendix A recalls them (coefficient-major SH, [xx,xy,xz,yy,yz,zz])."""
    stride = pod_stride(sh_fmt, cov_fmt)
    raw = np.frombuffer(np.ascontiguousarray(pods).view(np.uint8).tobytes(), dtype=np.uint8).reshape(n, stride)
    pos = raw[:, 0:12].copy().view(f32).reshape(n, 3)
    color = raw[:, 12:16].copy()
    q = raw[:, 16:16 + SH_BYTES[sh_fmt]]
    if sh_fmt == 0:
        sh = q.copy().view(f32).reshape(n, 45)
    elif sh_fmt == 1:
        sh = q.copy().view(np.float16).reshape(n, 46)[:, :45].astype(f32)
    elif sh_fmt == 2:
        mm = q[:, 0:4].copy().view(np.float16).astype(f32)
        t = q[:, 4:49].astype(f32) / f32(255.0)
        sh = mm[:, 0:1] * (f32(1.0) - t) + mm[:, 1:2] * t  # WGSL mix(min, max, t)
    else:
        sh = np.zeros((n, 45), dtype=f32)
    c = raw[:, 16 + SH_BYTES[sh_fmt]:16 + SH_BYTES[sh_fmt] + COV_BYTES[cov_fmt]]
    if cov_fmt == 0:
        cov = c.copy().view(f32).reshape(n, 6)
    elif cov_fmt == 1:
        cov = c.copy().view(np.float16).reshape(n, 6).astype(f32)
    else:
        raise NotImplementedError("rot+scale pods: cov3d is built in the shader; not restated in the twin")
    return pos, color, sh.astype(f32), cov


def fma32(a, b, c):
    """fmaf(a, b, c) on float32 arrays, exactly: the f64 product of two f32 is exact; the f64 sum is rounded once and its
    error recovered with TwoSum; if that sum sits exactly on an f32 rounding boundary the error's sign breaks the tie."""
    a = np.asarray(a, dtype=f32).astype(f64)
    b = np.asarray(b, dtype=f32).astype(f64)
    c = np.asarray(c, dtype=f32).astype(f64)
    with np.errstate(all="ignore"):
        p = a * b
        s = p + c
        bb = s - p
        err = (p - (s - bb)) + (c - bb)
        r = s.astype(f32)
        r64 = r.astype(f64)
        up = np.nextafter(r, f32(np.inf)).astype(f64)
        dn = np.nextafter(r, f32(-np.inf)).astype(f64)
        fix_up = (s - r64 == up - s) & (err > 0) & np.isfinite(s)
        fix_dn = (r64 - s == s - dn) & (err < 0) & np.isfinite(s)
        r = np.where(fix_up, up.astype(f32), np.where(fix_dn, dn.astype(f32), r))
    return r.astype(f32)


class DrawUniforms(Uniforms):
    """Adds what the vertex stage needs: camera position, S^-1 R^T, display mode / SH degree / no_sh0."""

    def __init__(self, cam, mt, gt):
        super().__init__(cam, mt, gt)
        rot = np.array(mt.rot, dtype=f32)
        scale = np.array(mt.scale, dtype=f32)
        r = _mat3_from_quat(rot)
        self.inv_sr = np.zeros((3, 3), dtype=f32)
        for c in range(3):
            for row in range(3):
                self.inv_sr[c, row] = r[row, c] / scale[row]
        v = self.view
        self.cam_pos = np.array([-((self.w[i, 0] * v[3, 0] + self.w[i, 1] * v[3, 1]) + self.w[i, 2] * v[3, 2]) for i in range(3)], dtype=f32)
        self.mode = int(gt.display_mode)
        self.sh_deg = int(gt.sh_deg)
        self.no_sh0 = bool(gt.no_sh0)


_SH_C1 = f32(0.4886025)
_SH_C2 = [f32(v) for v in (1.0925484, -1.0925484, 0.3153916, -1.0925484, 0.5462742)]
_SH_C3 = [f32(v) for v in (-0.5900436, 2.8906114, -0.4570458, 0.3731763, -0.4570458, 1.4453057, -0.5900436)]


def view_color_f32(u: DrawUniforms, color_u8, sh, dirs, sh_none=False):
    """utils.wesl:82-135 under the strict contract: basis factors individually rounded, each degree one fma chain."""
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    col = color_u8.astype(f32) / f32(255.0)
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    two, three, four = f32(2.0), f32(3.0), f32(4.0)
    b2 = [_SH_C2[0] * xy, _SH_C2[1] * yz, _SH_C2[2] * ((two * zz - xx) - yy), _SH_C2[3] * xz, _SH_C2[4] * (xx - yy)]
    b3 = [_SH_C3[0] * y * (three * xx - yy), _SH_C3[1] * xy * z, _SH_C3[2] * y * ((four * zz - xx) - yy),
          _SH_C3[3] * z * ((two * zz - three * xx) - three * yy), _SH_C3[4] * x * ((four * zz - xx) - yy),
          _SH_C3[5] * z * (xx - yy), _SH_C3[6] * x * (xx - three * yy)]
    out = np.zeros((len(x), 4), dtype=f32)
    for c in range(3):
        s = lambda i: sh[:, i * 3 + c]  # noqa: E731
        res = np.full(len(x), f32(0.5), dtype=f32) if u.no_sh0 else col[:, c].copy()
        if u.sh_deg >= 1 and not sh_none:
            t = fma32(s(1), z, -(s(0) * y))
            t = fma32(-s(2), x, t)
            res = fma32(_SH_C1, t, res)
            if u.sh_deg >= 2:
                acc = b2[0] * s(3)
                for k in range(1, 5):
                    acc = fma32(b2[k], s(3 + k), acc)
                res = res + acc
                if u.sh_deg >= 3:
                    acc = b3[0] * s(8)
                    for k in range(1, 7):
                        acc = fma32(b3[k], s(8 + k), acc)
                    res = res + acc
        out[:, c] = np.maximum(res, f32(0.0))
    out[:, 3] = col[:, 3]
    return out


def project_f32(pods, n, sh_fmt, cov_fmt, cam, mt, gt, indices):
    """vert_main once per instance (render.wesl:76-130) in the oracle's pixel-space form; dict of f32 arrays named like SoSplat."""
    u = DrawUniforms(cam, mt, gt)
    pos, color, sh, cov = unpack_pods(pods, n, sh_fmt, cov_fmt)
    idx = np.asarray(indices, dtype=np.int64)
    pos, color, sh, cov = pos[idx], color[idx], sh[idx], cov[idx]
    px, py, pz = pos[:, 0], pos[:, 1], pos[:, 2]
    m, pv = u.model, u.pv
    with np.errstate(all="ignore"):
        world = [((m[0, i] * px + m[1, i] * py) + m[2, i] * pz) + m[3, i] for i in range(3)]
        clip = [((pv[0, i] * world[0] + pv[1, i] * world[1]) + pv[2, i] * world[2]) + pv[3, i] for i in range(4)]
        nx, ny, nz = clip[0] / clip[3], clip[1] / clip[3], clip[2] / clip[3]
        out = {"z": nz.astype(f32)}
        out["cx"] = (nx + f32(1.0)) * f32(0.5) * u.size[0]
        out["cy"] = (f32(1.0) - ny) * f32(0.5) * u.size[1]
        vd = [u.cam_pos[i] - world[i] for i in range(3)]
        md = [(u.inv_sr[0, i] * vd[0] + u.inv_sr[1, i] * vd[1]) + u.inv_sr[2, i] * vd[2] for i in range(3)]
        inv_ml = f32(1.0) / np.sqrt((md[0] * md[0] + md[1] * md[1]) + md[2] * md[2])
        dirs = np.stack([-(md[0] * inv_ml), -(md[1] * inv_ml), -(md[2] * inv_ml)], axis=1).astype(f32)
        rgba = view_color_f32(u, color, sh, dirs, sh_none=(sh_fmt == 3))
        out["r"], out["g"], out["b"], out["a"] = rgba[:, 0], rgba[:, 1], rgba[:, 2], rgba[:, 3]
        if u.mode == 2:  # point: render.wesl:92-104
            vm = u.vm
            vp = [((vm[0, i] * px + vm[1, i] * py) + vm[2, i] * pz) + vm[3, i] for i in range(3)]
            ln = np.sqrt((vp[0] * vp[0] + vp[1] * vp[1]) + vp[2] * vp[2])
            half = f32(0.01) * u.gsize * f32(0.5) * u.size[1] / ln
            inv = f32(1.0) / half
            zero = np.zeros_like(inv)
            out.update(ax=inv, ay=zero, bx=zero, by=inv, ext_x=half, ext_y=half)
            out["valid"] = (half > 0) & np.isfinite(inv) & np.isfinite(out["cx"]) & np.isfinite(out["cy"])
            return out
        axes = cov2d_axes(u, pos, cov, u.std_dev * u.gsize)
        a0, a1, a2, a3 = axes[:, 0], axes[:, 1], axes[:, 2], axes[:, 3]
        mm = a0 * a0 + a1 * a1
        nn = a2 * a2 + a3 * a3
        two = f32(2.0)
        out["ax"] = two * a0 / mm
        out["ay"] = -(two * a1 / mm)
        out["bx"] = two * a2 / nn
        out["by"] = -(two * a3 / nn)
        hs = f32(0.5) * u.std_dev
        out["ext_x"] = hs * np.sqrt(a0 * a0 + a2 * a2)
        out["ext_y"] = hs * np.sqrt(a1 * a1 + a3 * a3)
        ok = np.ones(len(idx), dtype=bool)
        for k in ("ax", "ay", "bx", "by", "cx", "cy", "ext_x", "ext_y"):
            ok &= np.isfinite(out[k])
        out["valid"] = ok
    return out


def exp_neg_poly_f32(x):
    """so_exp_neg_poly: exp(-x), x >= 0, from exactly rounded steps (the 'strict' fragment exp of DESIGN.md §3)."""
    x = np.asarray(x, dtype=f32)
    y = x * f32(-1.44269504)
    n = np.rint(y).astype(f32)
    f = y - n
    p = np.full(x.shape, f32(1.54035304e-4), dtype=f32)
    for c in (1.33335581e-3, 9.61812911e-3, 5.55041087e-2, 2.40226507e-1, 6.93147181e-1, 1.0):
        p = fma32(p, f, f32(c))
    bits = p.view(np.uint32) + (n.astype(np.int32) << 23).astype(np.uint32)
    return np.where(n < f32(-125.0), f32(0.0), bits.view(f32)).astype(f32)


def composite_f32(sp, width, height, mode, std_dev, strict_exp=True):
    """frag_main + ALPHA_BLENDING into an rgba8unorm target cleared to BLACK, splats in the given order (strict contract):
    q = (fma(dx,ax,dy*ay), fma(dx,bx,dy*by)); r2 = fma(qx,qx,qy*qy); d <- rint(min(fma(d, 1-alpha, c255*alpha), 255))."""
    acc = np.zeros((height, width, 3), dtype=f32)
    sd = f32(std_dev)
    sd2 = sd * sd
    ol = (sd - f32(0.1)) * (sd - f32(0.1))
    n = len(sp["cx"])
    for k in range(n):
        if not sp["valid"][k]:
            continue
        cx, cy, ex, ey = sp["cx"][k], sp["cy"][k], sp["ext_x"][k], sp["ext_y"][k]
        x0, x1 = int(max(np.floor(cx - ex - 1.0), 0)), int(min(np.ceil(cx + ex + 1.0), width - 1))
        y0, y1 = int(max(np.floor(cy - ey - 1.0), 0)), int(min(np.ceil(cy + ey + 1.0), height - 1))
        if x1 < x0 or y1 < y0:
            continue
        dx = (np.arange(x0, x1 + 1, dtype=f32) + f32(0.5)) - cx
        dy = ((np.arange(y0, y1 + 1, dtype=f32) + f32(0.5)) - cy)[:, None]
        dxb = np.broadcast_to(dx, (dy.shape[0], dx.shape[0]))
        qx = fma32(dxb, sp["ax"][k], dy * sp["ay"][k])
        qy = fma32(dxb, sp["bx"][k], dy * sp["by"][k])
        a = sp["a"][k]
        if mode == 2:
            alive = (np.abs(qx) <= 1) & (np.abs(qy) <= 1)
            alpha = np.ones_like(qx)
        else:
            r2 = fma32(qx, qx, qy * qy)
            alive = r2 <= sd2
            if mode == 0:
                e = exp_neg_poly_f32(r2) if strict_exp else np.exp(-r2.astype(f64)).astype(f32)
                alpha = (a * e).astype(f32)
            else:
                alpha = (a + (f32(1.0) - a) * (r2 > ol).astype(f32)).astype(f32)
        if not alive.any():
            continue
        om = (f32(1.0) - alpha).astype(f32)
        win = acc[y0:y1 + 1, x0:x1 + 1]
        for c, name in enumerate(("r", "g", "b")):
            c255 = np.minimum(sp[name][k] * f32(255.0), f32(255.0))
            new = np.rint(np.minimum(fma32(win[..., c], om, c255 * alpha), f32(255.0)))
            win[..., c] = np.where(alive, new, win[..., c])
    img = np.zeros((height, width, 4), dtype=np.uint8)
    img[..., :3] = acc.astype(np.uint8)
    img[..., 3] = 255
    return img


def draw_reference_f64(pods, n, sh_fmt, cov_fmt, cam, mt, gt, order, width, height):
    """The draw stage from the WGSL text, float64, shader formulation (see the section header).  rgba8unorm, cleared BLACK."""
    pos, color, sh, cov = unpack_pods(pods, n, sh_fmt, cov_fmt)
    V = np.array(cam.view, dtype=f64).reshape(4, 4).T          # row-major maths below
    P = np.array(cam.proj, dtype=f64).reshape(4, 4).T
    size = np.array(cam.size, dtype=f64)
    q = np.array(mt.rot, dtype=f64)
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    S = np.diag(np.array(mt.scale, dtype=f64))
    SR = R @ S                                                  # model_scale_rot_mat
    M = np.eye(4)
    M[:3, :3] = SR
    M[:3, 3] = np.array(mt.pos, dtype=f64)
    inv_sr = np.linalg.inv(S) @ R.T                             # model_transform_inv_sr_mat
    W3 = V[:3, :3]
    cam_pos = -(W3.T @ V[:3, 3])                                # render.wesl:59-63
    sd = float(gt.max_std_dev) / 255.0 * 3.0
    gsize = float(gt.size)
    mode, sh_deg, no_sh0 = int(gt.display_mode), int(gt.sh_deg), bool(gt.no_sh0)
    focal = np.array([P[0, 0], P[1, 1]]) * size * 0.5
    acc = np.zeros((height, width, 3), dtype=f64)
    c1 = 0.4886025
    c2 = (1.0925484, -1.0925484, 0.3153916, -1.0925484, 0.5462742)
    c3 = (-0.5900436, 2.8906114, -0.4570458, 0.3731763, -0.4570458, 1.4453057, -0.5900436)
    for g in np.asarray(order, dtype=np.int64):
        p = pos[g].astype(f64)
        world = M @ np.append(p, 1.0)
        view_pos = V @ world
        proj = P @ view_pos
        if not (proj[3] > 0.0 and 0.0 <= proj[2] <= proj[3]):   # whole quad clipped (flat z, w: render.wesl:123)
            continue
        # colour
        d = inv_sr @ (cam_pos - world[:3])
        d = -d / np.linalg.norm(d)
        dx_, dy_, dz_ = d
        col = color[g].astype(f64) / 255.0
        res = np.full(3, 0.5) if no_sh0 else col[:3].copy()
        s = sh[g].astype(f64).reshape(15, 3)
        if sh_deg >= 1 and sh_fmt != 3:
            res = res + c1 * (-s[0] * dy_ + s[1] * dz_ - s[2] * dx_)
            if sh_deg >= 2:
                xx, yy, zz, xy, yz, xz = dx_ * dx_, dy_ * dy_, dz_ * dz_, dx_ * dy_, dy_ * dz_, dx_ * dz_
                res = res + (c2[0] * xy * s[3] + c2[1] * yz * s[4] + c2[2] * (2 * zz - xx - yy) * s[5] + c2[3] * xz * s[6]
                             + c2[4] * (xx - yy) * s[7])
                if sh_deg >= 3:
                    res = res + (c3[0] * dy_ * (3 * xx - yy) * s[8] + c3[1] * xy * dz_ * s[9] + c3[2] * dy_ * (4 * zz - xx - yy) * s[10]
                                 + c3[3] * dz_ * (2 * zz - 3 * xx - 3 * yy) * s[11] + c3[4] * dx_ * (4 * zz - xx - yy) * s[12]
                                 + c3[5] * dz_ * (xx - yy) * s[13] + c3[6] * dx_ * (xx - 3 * yy) * s[14])
        rgb = np.maximum(res, 0.0)
        alpha0 = col[3]
        ndc = proj[:2] / proj[3]
        centre = np.array([(ndc[0] + 1.0) * 0.5 * size[0], (1.0 - ndc[1]) * 0.5 * size[1]])
        if mode == 2:
            half_px = 0.01 * gsize * (size[1] / size[0]) / np.linalg.norm(view_pos[:3]) * size[0] * 0.5
            A = np.array([[half_px, 0.0], [0.0, -half_px]])     # pixel delta per unit corner offset (corners at +-1)
            lim = 1.0
        else:
            vrk = np.array([[cov[g, 0], cov[g, 1], cov[g, 2]], [cov[g, 1], cov[g, 3], cov[g, 4]], [cov[g, 2], cov[g, 4], cov[g, 5]]], dtype=f64)
            t = V @ M @ np.append(p, 1.0)
            J = np.array([[focal[0] / t[2], 0.0, -(focal[0] * t[0]) / (t[2] * t[2])],
                          [0.0, focal[1] / t[2], -(focal[1] * t[1]) / (t[2] * t[2])],
                          [0.0, 0.0, 0.0]])
            T = J @ W3 @ SR
            c2d = T @ vrk @ T.T
            cxx, cxy, cyy = c2d[0, 0], c2d[0, 1], c2d[1, 1]
            mid = 0.5 * (cxx + cyy)
            radius = np.hypot(0.5 * (cxx - cyy), cxy)
            lam1, lam2 = mid + radius, mid - radius
            if lam2 < 0.0 or not np.isfinite(lam1):
                continue
            diag = np.array([cxy, lam1 - cxx])
            dd = np.array([0.0, 1.0]) if (diag[0] == 0.0 and diag[1] == 0.0) else diag / np.linalg.norm(diag)
            major = min(sd * gsize * np.sqrt(lam1), 1024.0) * dd
            minor = min(sd * gsize * np.sqrt(lam2), 1024.0) * np.array([dd[1], -dd[0]])
            if not (major.any() or minor.any()):
                continue
            # clip offset = qo.x * w * major / size + qo.y * w * minor / size  ->  pixel delta = A @ qo
            A = 0.5 * np.array([[major[0], minor[0]], [-major[1], -minor[1]]])
            lim = sd
        if abs(np.linalg.det(A)) < 1e-300:
            continue
        Ainv = np.linalg.inv(A)
        ext = np.abs(A) @ np.array([lim, lim])
        x0, x1 = int(max(np.floor(centre[0] - ext[0] - 1), 0)), int(min(np.ceil(centre[0] + ext[0] + 1), width - 1))
        y0, y1 = int(max(np.floor(centre[1] - ext[1] - 1), 0)), int(min(np.ceil(centre[1] + ext[1] + 1), height - 1))
        if x1 < x0 or y1 < y0:
            continue
        px = (np.arange(x0, x1 + 1) + 0.5) - centre[0]
        py = ((np.arange(y0, y1 + 1) + 0.5) - centre[1])[:, None]
        qx = Ainv[0, 0] * px + Ainv[0, 1] * py
        qy = Ainv[1, 0] * px + Ainv[1, 1] * py
        inside = (np.abs(qx) <= lim) & (np.abs(qy) <= lim)
        if mode == 2:
            alive, alpha = inside, np.ones_like(qx)
        else:
            r2 = qx * qx + qy * qy
            alive = inside & (r2 <= sd * sd)
            if mode == 0:
                alpha = alpha0 * np.exp(-r2)
            else:
                alpha = alpha0 + (1.0 - alpha0) * (r2 > (sd - 0.1) ** 2)
        if not alive.any():
            continue
        src = np.minimum(rgb, 1.0) * 255.0                     # fixed-point attachment: source clamped to [0,1]
        win = acc[y0:y1 + 1, x0:x1 + 1]
        new = np.rint(np.minimum(win * (1.0 - alpha)[..., None] + src[None, None, :] * alpha[..., None], 255.0))
        win[...] = np.where(alive[..., None], new, win)
    img = np.zeros((height, width, 4), dtype=np.uint8)
    img[..., :3] = acc.astype(np.uint8)
    img[..., 3] = 255
    return img
