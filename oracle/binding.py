"""ctypes binding of the CPU oracle (oracle/splat_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsplat_oracle.so")

GAUSSIAN_DTYPE = np.dtype(
    [("pos", "<f4", 3), ("color", "u1", 4), ("sh", "<f4", 45), ("scale", "<f4", 3), ("rot", "<f4", 4)]
)
assert GAUSSIAN_DTYPE.itemsize == 224

SH_SINGLE, SH_HALF, SH_NORM8, SH_NONE = 0, 1, 2, 3
COV_SINGLE, COV_HALF, COV_ROT_SCALE = 0, 1, 2
MODE_SPLAT, MODE_ELLIPSE, MODE_POINT = 0, 1, 2
TARGET_RGBA8, TARGET_BGRA8, TARGET_RGBA16F, TARGET_RGBA32F, TARGET_RGBA8_SRGB, TARGET_BGRA8_SRGB = 0, 1, 2, 3, 4, 5


class CameraPod(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("size", C.c_float * 2), ("pad", C.c_uint32 * 2)]


class ModelTransformPod(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("pad0", C.c_float), ("rot", C.c_float * 4), ("scale", C.c_float * 3), ("pad1", C.c_float)]


class GaussianTransformPod(C.Structure):
    _fields_ = [("size", C.c_float), ("display_mode", C.c_uint8), ("sh_deg", C.c_uint8), ("no_sh0", C.c_uint8), ("max_std_dev", C.c_uint8)]


class Model(C.Structure):
    _fields_ = [
        ("pods", C.c_void_p), ("n", C.c_uint32), ("sh_fmt", C.c_int32), ("cov_fmt", C.c_int32),
        ("model_transform", ModelTransformPod), ("selection", C.c_void_p), ("invert_selection", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [("visible", C.c_uint64), ("bbox_pixels", C.c_uint64), ("alive_pixels", C.c_uint64)]


SPLAT_DTYPE = np.dtype(
    [("cx", "<f4"), ("cy", "<f4"), ("ax", "<f4"), ("ay", "<f4"), ("bx", "<f4"), ("by", "<f4"),
     ("r", "<f4"), ("g", "<f4"), ("b", "<f4"), ("a", "<f4"), ("ext_x", "<f4"), ("ext_y", "<f4"), ("valid", "<i4"), ("z", "<f4")]
)

assert C.sizeof(CameraPod) == 144 and C.sizeof(ModelTransformPod) == 48 and C.sizeof(GaussianTransformPod) == 8


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) into oracle/_build/."""
    src = os.path.join(_HERE, "splat_oracle.c")
    hdr = os.path.join(_HERE, "splat_oracle.h")
    tab = os.path.join(_HERE, "srgb_tables.h")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(tab)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
         "-o", _SO, src, "-lm"]
    )
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.so_pod_stride.restype = C.c_uint32
        l.so_pod_stride.argtypes = [C.c_int, C.c_int]
        l.so_pack_gaussians.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        l.so_gaussians_from_ply_props.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        l.so_camera_pod.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(CameraPod)]
        l.so_model_transform_pod.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ModelTransformPod)]
        l.so_gaussian_transform_pod.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(GaussianTransformPod)]
        l.so_preprocess.restype = C.c_uint32
        l.so_preprocess.argtypes = [C.POINTER(Model), C.POINTER(CameraPod), C.POINTER(GaussianTransformPod),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.so_radix_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        l.so_project.argtypes = [C.POINTER(Model), C.POINTER(CameraPod), C.POINTER(GaussianTransformPod), C.c_void_p, C.c_uint32, C.c_void_p]
        l.so_render.argtypes = [C.POINTER(Model), C.c_uint32, C.POINTER(CameraPod), C.POINTER(GaussianTransformPod),
                                C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(Stats), C.c_int]
        l.so_exp_neg_poly.restype = C.c_float
        l.so_exp_neg_poly.argtypes = [C.c_float]
        l.so_select_rect.argtypes = [C.POINTER(Model), C.POINTER(CameraPod), C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        l.so_render_pass.argtypes = [C.POINTER(Model), C.POINTER(CameraPod), C.POINTER(GaussianTransformPod), C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.so_draw.argtypes = [C.POINTER(Model), C.POINTER(CameraPod), C.POINTER(GaussianTransformPod), C.c_void_p, C.c_uint32, C.c_int,
                              C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.so_select_brush.argtypes = [C.POINTER(Model), C.POINTER(CameraPod), C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_void_p]
        l.so_max_threads.restype = C.c_int
        l.so_set_threads.argtypes = [C.c_int]
        _lib = l
    return _lib


def _f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


def pod_stride(sh_fmt: int, cov_fmt: int) -> int:
    return lib().so_pod_stride(sh_fmt, cov_fmt)


def pack_gaussians(gaussians: np.ndarray, sh_fmt: int = SH_SINGLE, cov_fmt: int = COV_SINGLE) -> np.ndarray:
    g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN_DTYPE)
    out = np.zeros(len(g) * pod_stride(sh_fmt, cov_fmt), dtype=np.uint8)
    lib().so_pack_gaussians(g.ctypes.data, len(g), sh_fmt, cov_fmt, out.ctypes.data)
    return out


def gaussians_from_ply_props(props: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(props, dtype=np.float32).reshape(-1, 62)
    out = np.zeros(len(p), dtype=GAUSSIAN_DTYPE)
    lib().so_gaussians_from_ply_props(p.ctypes.data, len(p), out.ctypes.data)
    return out


def camera_pod(pos, yaw, pitch, width, height, z_near=0.1, z_far=1e4, fov_y=np.float32(np.deg2rad(np.float32(60.0)))) -> CameraPod:
    pod = CameraPod()
    p = _f3(pos)
    lib().so_camera_pod(p.ctypes.data, yaw, pitch, z_near, z_far, float(fov_y), width, height, C.byref(pod))
    return pod


def model_transform_pod(pos=(0, 0, 0), rot=(0, 0, 0, 1), scale=(1, 1, 1)) -> ModelTransformPod:
    pod = ModelTransformPod()
    a, b, c = _f3(pos), _f3(rot), _f3(scale)
    lib().so_model_transform_pod(a.ctypes.data, b.ctypes.data, c.ctypes.data, C.byref(pod))
    return pod


def gaussian_transform_pod(size=1.0, display_mode=MODE_SPLAT, sh_deg=3, no_sh0=False, max_std_dev=3.0) -> GaussianTransformPod:
    pod = GaussianTransformPod()
    lib().so_gaussian_transform_pod(size, display_mode, sh_deg, int(no_sh0), max_std_dev, C.byref(pod))
    return pod


class OracleModel:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, pods: np.ndarray, n: int, sh_fmt=SH_SINGLE, cov_fmt=COV_SINGLE,
                 model_transform: ModelTransformPod | None = None, selection: np.ndarray | None = None,
                 invert_selection: int = 1):
        self.pods = np.ascontiguousarray(pods, dtype=np.uint8)
        assert self.pods.size == n * pod_stride(sh_fmt, cov_fmt)
        self.selection = None if selection is None else np.ascontiguousarray(selection, dtype=np.uint32)
        self.c = Model()
        self.c.pods = self.pods.ctypes.data
        self.c.n = n
        self.c.sh_fmt, self.c.cov_fmt = sh_fmt, cov_fmt
        self.c.model_transform = model_transform if model_transform is not None else model_transform_pod()
        self.c.selection = None if self.selection is None else self.selection.ctypes.data
        self.c.invert_selection = invert_selection
        self.n = n


def padded_keys(n: int) -> int:
    return (n + 3839) // 3840 * 3840


def preprocess(model: OracleModel, cam: CameraPod, gt: GaussianTransformPod):
    n = model.n
    cap = max(padded_keys(n), 1)
    idx = np.zeros(cap, dtype=np.uint32)
    keys = np.zeros(cap, dtype=np.float32)
    mask = np.zeros(max((n + 31) // 32, 1), dtype=np.uint32)
    draw = np.zeros(4, dtype=np.uint32)
    sort = np.zeros(3, dtype=np.uint32)
    v = lib().so_preprocess(C.byref(model.c), C.byref(cam), C.byref(gt), idx.ctypes.data, keys.ctypes.data,
                            mask.ctypes.data, draw.ctypes.data, sort.ctypes.data)
    return dict(count=int(v), indices=idx, keys=keys, mask=mask[: (n + 31) // 32], draw_args=draw, sort_args=sort)


def radix_sort(keys_u32: np.ndarray, payload: np.ndarray, count: int | None = None):
    k = np.ascontiguousarray(keys_u32).view(np.uint32).copy()
    p = np.ascontiguousarray(payload, dtype=np.uint32).copy()
    lib().so_radix_sort(k.ctypes.data, p.ctypes.data, len(k) if count is None else count)
    return k, p


def project(model: OracleModel, cam: CameraPod, gt: GaussianTransformPod, indices: np.ndarray) -> np.ndarray:
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    out = np.zeros(len(idx), dtype=SPLAT_DTYPE)
    lib().so_project(C.byref(model.c), C.byref(cam), C.byref(gt), idx.ctypes.data, len(idx), out.ctypes.data)
    return out


def render(models, cam: CameraPod, gt: GaussianTransformPod, target_format=TARGET_RGBA8, strict_exp=False,
           row0=0, rows=None, n_threads=0):
    if isinstance(models, OracleModel):
        models = [models]
    w, h = int(cam.size[0]), int(cam.size[1])
    rows = h - row0 if rows is None else rows
    arr = (Model * len(models))(*[m.c for m in models])
    dt = {TARGET_RGBA8: np.uint8, TARGET_BGRA8: np.uint8, TARGET_RGBA16F: np.uint16, TARGET_RGBA32F: np.float32,
          TARGET_RGBA8_SRGB: np.uint8, TARGET_BGRA8_SRGB: np.uint8}[target_format]
    out = np.zeros((rows, w, 4), dtype=dt)
    st = Stats()
    lib().so_render(arr, len(models), C.byref(cam), C.byref(gt), target_format, int(strict_exp), row0, rows,
                    out.ctypes.data, C.byref(st), n_threads)
    if target_format == TARGET_RGBA16F:
        out = out.view(np.float16)
    return out, dict(visible=st.visible, bbox_pixels=st.bbox_pixels, alive_pixels=st.alive_pixels)


def render_pass(model: OracleModel, cam: CameraPod, gt: GaussianTransformPod, target: np.ndarray, load: bool, depth: np.ndarray | None = None,
                compare: int = 8, depth_write: bool = False, target_format=TARGET_RGBA8, strict_exp=False, n_threads=0):
    """Renderer::render_with_pass: composites `model` into `target` (in place; loaded or cleared) against the optional
    f32 depth attachment `depth` (updated in place when depth_write)."""
    assert target.flags["C_CONTIGUOUS"] and (depth is None or (depth.dtype == np.float32 and depth.flags["C_CONTIGUOUS"]))
    buf = target.view(np.uint16) if target.dtype == np.float16 else target
    lib().so_render_pass(C.byref(model.c), C.byref(cam), C.byref(gt), target_format, int(strict_exp), int(load), buf.ctypes.data,
                         None if depth is None else depth.ctypes.data, compare, int(depth_write), n_threads)
    return target


def draw(model: OracleModel, cam: CameraPod, gt: GaussianTransformPod, indices: np.ndarray, target: np.ndarray, load: bool = False,
         depth: np.ndarray | None = None, compare: int = 8, depth_write: bool = False, target_format=TARGET_RGBA8, strict_exp=False,
         n_threads=0):
    """Renderer<G, ()>::render on a caller's indices: instance i = Gaussian indices[i], drawn in that order into `target`
    (in place; loaded or cleared), no preprocess and no sort."""
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    assert target.flags["C_CONTIGUOUS"] and (depth is None or (depth.dtype == np.float32 and depth.flags["C_CONTIGUOUS"]))
    buf = target.view(np.uint16) if target.dtype == np.float16 else target
    lib().so_draw(C.byref(model.c), C.byref(cam), C.byref(gt), idx.ctypes.data, len(idx), target_format, int(strict_exp), int(load),
                  buf.ctypes.data, None if depth is None else depth.ctypes.data, compare, int(depth_write), n_threads)
    return target


def select_rect(model: OracleModel, cam: CameraPod, x0, y0, x1, y1) -> np.ndarray:
    out = np.zeros(max((model.n + 31) // 32, 1), dtype=np.uint32)
    lib().so_select_rect(C.byref(model.c), C.byref(cam), x0, y0, x1, y1, out.ctypes.data)
    return out[: (model.n + 31) // 32]


def select_brush(model: OracleModel, cam: CameraPod, points, radius, accumulate=False, dest: np.ndarray | None = None) -> np.ndarray:
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
    words = (model.n + 31) // 32
    out = np.zeros(max(words, 1), dtype=np.uint32)
    if dest is not None:
        out[:words] = dest[:words]
    lib().so_select_brush(C.byref(model.c), C.byref(cam), pts.ctypes.data, len(pts), radius, int(accumulate), out.ctypes.data)
    return out[:words]


def exp_neg_poly(x: float) -> float:
    return lib().so_exp_neg_poly(x)


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use every core the process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().so_set_threads(n)
    return lib().so_max_threads()
