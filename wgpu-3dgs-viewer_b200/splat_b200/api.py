"""ctypes binding of include/splat_b200.h — one Python method per C entry point."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# SB_LIB: load a tuning variant built next to the default library (Makefile EXTRA/BUILD/LIB); never a fallback
_LIB = os.environ.get("SB_LIB") or os.path.join(_PKG, "lib", "libsplat_b200.so")

SH_SINGLE, SH_HALF, SH_NORM8, SH_NONE = 0, 1, 2, 3
COV_SINGLE, COV_HALF, COV_ROT_SCALE = 0, 1, 2
MODE_SPLAT, MODE_ELLIPSE, MODE_POINT = 0, 1, 2
TARGET_RGBA8, TARGET_BGRA8, TARGET_RGBA16F, TARGET_RGBA32F, TARGET_RGBA8_SRGB, TARGET_BGRA8_SRGB = 0, 1, 2, 3, 4, 5

GAUSSIAN_DTYPE = np.dtype(
    [("pos", "<f4", 3), ("color", "u1", 4), ("sh", "<f4", 45), ("scale", "<f4", 3), ("rot", "<f4", 4)]
)

# every symbol include/splat_b200.h declares (checked by tests/test_abi.py against the .so)
EXPORTED_SYMBOLS = [
    "sb_version", "sb_status_string", "sb_pod_stride", "sb_pack_gaussians", "sb_keys_buffer_size_bytes",
    "sb_padded_key_count", "sb_camera_pod", "sb_model_transform_pod", "sb_gaussian_transform_pod", "sb_read_ply",
    "sb_free", "sb_ctx_create", "sb_ctx_destroy", "sb_last_error_string", "sb_ctx_set_model_size_limit",
    "sb_viewer_create", "sb_viewer_create_from_gaussians", "sb_viewer_create_from_device", "sb_viewer_destroy",
    "sb_viewer_update_camera_with_pod", "sb_viewer_update_camera", "sb_viewer_update_model_transform",
    "sb_viewer_update_model_transform_with_pod", "sb_viewer_update_gaussian_transform",
    "sb_viewer_update_gaussian_transform_with_pod", "sb_viewer_enable_selection", "sb_viewer_selection_ptr",
    "sb_viewer_set_selection", "sb_viewer_read_selection", "sb_viewer_set_invert_selection", "sb_viewer_select_rect",
    "sb_viewer_select_brush", "sb_viewer_apply_rgb_override", "sb_viewer_restore_gaussians", "sb_viewer_render", "sb_viewer_render_with_pass",
    "sb_viewer_preprocess", "sb_viewer_sort", "sb_viewer_draw", "sb_viewer_render_to_host", "sb_viewer_render_batch", "sb_viewer_gaussians_ptr",
    "sb_viewer_indirect_args_ptr", "sb_viewer_radix_sort_indirect_args_ptr", "sb_viewer_indirect_indices_ptr",
    "sb_viewer_gaussians_depth_ptr", "sb_viewer_read_indirect_args", "sb_viewer_read_indices",
    "sb_viewer_read_depth_keys", "sb_viewer_read_frame_stats", "sb_viewer_raster_path", "sb_viewer_set_strict_exp",
    "sb_viewer_set_exact_cutoff",
    "sb_viewer_set_stage_timing", "sb_viewer_read_stage_times", "sb_viewer_set_raster_counting",
    "sb_viewer_read_raster_counters",
    "sb_viewer_read_raster_warp_counters",
    "sb_viewer_reserve_duplicates", "sb_sorter_create", "sb_sorter_destroy", "sb_sorter_sort",
    "sb_preprocessor_create", "sb_preprocessor_destroy", "sb_preprocessor_preprocess",
    "sb_renderer_create", "sb_renderer_destroy", "sb_renderer_render", "sb_renderer_set_strict_exp", "sb_mm_create",
    "sb_mm_destroy", "sb_mm_insert_model", "sb_mm_remove_model", "sb_mm_update_camera_with_pod",
    "sb_mm_update_model_transform_with_pod", "sb_mm_update_gaussian_transform_with_pod", "sb_mm_set_selection",
    "sb_mm_render", "sb_mm_read_model_indices",
    "sb_mm_update_camera", "sb_mm_update_model_transform", "sb_mm_update_gaussian_transform", "sb_mm_insert_model_from_gaussians",
    "sb_mm_insert_model_from_device", "sb_mm_select_rect", "sb_mm_select_brush", "sb_mm_enable_selection", "sb_mm_read_selection",
    "sb_mm_render_with_pass",
    "sb_read_spz", "sb_viewer_apply_basic_color_modifiers", "sb_strips_create", "sb_strips_connect", "sb_strips_scatter", "sb_strips_render",
    "sb_strips_destroy", "sb_probe_peaks", "sb_viewer_set_strip_cull", "sb_viewer_read_tile_row_work", "sb_shared_frame_create", "sb_shared_frame_open", "sb_shared_frame_close", "sb_shared_frame_destroy",
]


class CameraPod(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("size", C.c_float * 2), ("_padding", C.c_uint32 * 2)]


class ModelTransformPod(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("_pad0", C.c_float), ("rot", C.c_float * 4), ("scale", C.c_float * 3), ("_pad1", C.c_float)]


class GaussianTransformPod(C.Structure):
    _fields_ = [("size", C.c_float), ("display_mode", C.c_uint8), ("sh_deg", C.c_uint8), ("no_sh0", C.c_uint8), ("max_std_dev", C.c_uint8)]


class BasicColorModifiers(C.Structure):
    _fields_ = [("rgb_override", C.c_int32), ("rgb_or_hsv", C.c_float * 3), ("alpha", C.c_float), ("contrast", C.c_float),
                ("exposure", C.c_float), ("gamma", C.c_float)]


class DrawIndirectArgs(C.Structure):
    _fields_ = [("vertex_count", C.c_uint32), ("instance_count", C.c_uint32), ("first_vertex", C.c_uint32), ("first_instance", C.c_uint32)]


class DispatchIndirectArgs(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32)]


class Target(C.Structure):
    _fields_ = [("d_pixels", C.c_void_p), ("pitch_bytes", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("format", C.c_int32), ("row0", C.c_uint32), ("rows", C.c_uint32)]


class DepthAttachment(C.Structure):
    _fields_ = [("d_depth", C.c_void_p), ("pitch_bytes", C.c_uint32), ("compare", C.c_int32), ("write_enabled", C.c_int32)]


class PreprocessorBindGroup(C.Structure):
    _fields_ = [("camera", CameraPod), ("model_transform", ModelTransformPod), ("gaussian_transform", GaussianTransformPod),
                ("d_gaussians", C.c_void_p), ("gaussians_bytes", C.c_uint64), ("d_indirect_args", C.c_void_p),
                ("d_radix_sort_indirect_args", C.c_void_p), ("d_indirect_indices", C.c_void_p), ("indirect_indices_bytes", C.c_uint64),
                ("d_gaussians_depth", C.c_void_p), ("gaussians_depth_bytes", C.c_uint64), ("d_selection", C.c_void_p),
                ("invert_selection", C.c_uint32)]


class RendererBindGroup(C.Structure):
    _fields_ = [("camera", CameraPod), ("model_transform", ModelTransformPod), ("gaussian_transform", GaussianTransformPod),
                ("d_gaussians", C.c_void_p), ("gaussians_bytes", C.c_uint64), ("d_indirect_indices", C.c_void_p),
                ("indirect_indices_bytes", C.c_uint64)]


assert C.sizeof(CameraPod) == 144 and C.sizeof(ModelTransformPod) == 48 and C.sizeof(GaussianTransformPod) == 8
assert C.sizeof(DrawIndirectArgs) == 16 and C.sizeof(DispatchIndirectArgs) == 12


class SplatError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"splat_b200 status {status}: {message}")
        self.status = status


def lib_path() -> str:
    return _LIB


def build(verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", _PKG, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libsplat_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return _LIB


_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library.  Fails loudly when it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise ImportError(f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (make -C {_PKG})")
    l = C.CDLL(_LIB)
    vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
    P = C.POINTER

    def sig(name, res, *args):
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("sb_version", C.c_char_p)
    sig("sb_status_string", C.c_char_p, i32)
    sig("sb_pod_stride", u32, i32, i32)
    sig("sb_pack_gaussians", i32, vp, u64, i32, i32, vp)
    sig("sb_keys_buffer_size_bytes", u64, u32)
    sig("sb_padded_key_count", u32, u32)
    sig("sb_camera_pod", i32, vp, f32, f32, f32, f32, f32, u32, u32, P(CameraPod))
    sig("sb_model_transform_pod", i32, vp, vp, vp, P(ModelTransformPod))
    sig("sb_gaussian_transform_pod", i32, f32, i32, i32, i32, f32, P(GaussianTransformPod))
    sig("sb_read_ply", i32, C.c_char_p, P(vp), P(u64))
    sig("sb_read_spz", i32, C.c_char_p, P(vp), P(u64))
    sig("sb_viewer_apply_basic_color_modifiers", i32, vp, vp, P(BasicColorModifiers))
    sig("sb_free", None, vp)
    sig("sb_ctx_create", i32, i32, P(vp))
    sig("sb_ctx_destroy", None, vp)
    sig("sb_last_error_string", C.c_char_p, vp)
    sig("sb_ctx_set_model_size_limit", i32, vp, u64)
    sig("sb_viewer_create", i32, vp, i32, i32, i32, vp, u64, P(vp))
    sig("sb_viewer_create_from_gaussians", i32, vp, i32, i32, i32, vp, u64, P(vp))
    sig("sb_viewer_create_from_device", i32, vp, i32, i32, i32, vp, u64, u64, P(vp))
    sig("sb_viewer_destroy", None, vp)
    sig("sb_viewer_update_camera_with_pod", i32, vp, P(CameraPod))
    sig("sb_viewer_update_camera", i32, vp, vp, f32, f32, f32, f32, f32, u32, u32)
    sig("sb_viewer_update_model_transform", i32, vp, vp, vp, vp)
    sig("sb_viewer_update_model_transform_with_pod", i32, vp, P(ModelTransformPod))
    sig("sb_viewer_update_gaussian_transform", i32, vp, f32, i32, i32, i32, f32)
    sig("sb_viewer_update_gaussian_transform_with_pod", i32, vp, P(GaussianTransformPod))
    sig("sb_viewer_enable_selection", i32, vp, i32)
    sig("sb_viewer_selection_ptr", i32, vp, P(vp), P(u64))
    sig("sb_viewer_set_selection", i32, vp, vp, vp, u64)
    sig("sb_viewer_read_selection", i32, vp, vp, vp, u64)
    sig("sb_viewer_set_invert_selection", i32, vp, i32)
    sig("sb_viewer_select_rect", i32, vp, vp, f32, f32, f32, f32)
    sig("sb_viewer_select_brush", i32, vp, vp, P(f32), u32, f32, i32)
    sig("sb_viewer_apply_rgb_override", i32, vp, vp, P(f32), f32)
    sig("sb_viewer_restore_gaussians", i32, vp, vp)
    sig("sb_viewer_render_with_pass", i32, vp, vp, P(Target), P(DepthAttachment), i32, i32)
    sig("sb_viewer_render", i32, vp, vp, P(Target))
    sig("sb_viewer_preprocess", i32, vp, vp)
    sig("sb_viewer_sort", i32, vp, vp)
    sig("sb_viewer_draw", i32, vp, vp, P(Target))
    sig("sb_viewer_render_to_host", i32, vp, vp, P(CameraPod), vp, u64)
    sig("sb_viewer_render_batch", i32, vp, vp, P(CameraPod), P(Target), P(vp), u32)
    sig("sb_viewer_gaussians_ptr", i32, vp, P(vp), P(u64))
    sig("sb_viewer_indirect_args_ptr", i32, vp, P(vp))
    sig("sb_viewer_radix_sort_indirect_args_ptr", i32, vp, P(vp))
    sig("sb_viewer_indirect_indices_ptr", i32, vp, P(vp), P(u64))
    sig("sb_viewer_gaussians_depth_ptr", i32, vp, P(vp), P(u64))
    sig("sb_viewer_read_indirect_args", i32, vp, vp, P(DrawIndirectArgs), P(DispatchIndirectArgs))
    sig("sb_viewer_read_indices", i32, vp, vp, vp, u64)
    sig("sb_viewer_read_depth_keys", i32, vp, vp, vp, u64)
    sig("sb_viewer_read_frame_stats", i32, vp, vp, P(u64), P(u64), P(u32))
    sig("sb_viewer_raster_path", i32, vp, P(i32))
    sig("sb_viewer_set_strict_exp", i32, vp, i32)
    sig("sb_viewer_set_exact_cutoff", i32, vp, i32)
    sig("sb_viewer_set_strip_cull", i32, vp, i32)
    sig("sb_probe_peaks", i32, vp, vp, P(C.c_double), P(C.c_double))
    sig("sb_viewer_read_tile_row_work", i32, vp, vp, P(u64), u32)
    sig("sb_shared_frame_create", i32, vp, u64, P(vp), C.c_char_p)
    sig("sb_shared_frame_open", i32, vp, C.c_char_p, P(vp))
    sig("sb_shared_frame_close", i32, vp, vp)
    sig("sb_shared_frame_destroy", i32, vp, vp)
    sig("sb_strips_create", i32, vp, u32, u32, P(u32), P(u32), P(vp), C.c_char_p)
    sig("sb_strips_connect", i32, vp, C.c_char_p)
    sig("sb_strips_scatter", i32, vp, vp)
    sig("sb_strips_render", i32, vp, vp, P(Target))
    sig("sb_strips_destroy", None, vp)
    sig("sb_viewer_set_raster_counting", i32, vp, i32)
    sig("sb_viewer_read_raster_counters", i32, vp, vp, P(u64), P(u64))
    sig("sb_viewer_read_raster_warp_counters", i32, vp, vp, P(u64), P(u64))
    sig("sb_viewer_set_stage_timing", i32, vp, i32)
    sig("sb_viewer_read_stage_times", i32, vp, vp, P(f32))
    sig("sb_viewer_reserve_duplicates", i32, vp, u64)
    sig("sb_sorter_create", i32, vp, u32, P(vp))
    sig("sb_sorter_destroy", None, vp)
    sig("sb_sorter_sort", i32, vp, vp, vp, vp, vp, u32, i32, i32)
    sig("sb_preprocessor_create", i32, vp, i32, i32, u64, P(vp))
    sig("sb_preprocessor_destroy", None, vp)
    sig("sb_preprocessor_preprocess", i32, vp, vp, P(PreprocessorBindGroup), u32)
    sig("sb_renderer_create", i32, vp, i32, i32, i32, u64, P(vp))
    sig("sb_renderer_destroy", None, vp)
    sig("sb_renderer_render", i32, vp, vp, P(RendererBindGroup), P(Target), vp, P(DepthAttachment), i32)
    sig("sb_renderer_set_strict_exp", i32, vp, i32)
    sig("sb_mm_create", i32, vp, i32, i32, i32, P(vp))
    sig("sb_mm_destroy", None, vp)
    sig("sb_mm_insert_model", i32, vp, u64, vp, u64, P(i32))
    sig("sb_mm_remove_model", i32, vp, u64, P(i32))
    sig("sb_mm_update_camera_with_pod", i32, vp, P(CameraPod))
    sig("sb_mm_update_model_transform_with_pod", i32, vp, u64, P(ModelTransformPod))
    sig("sb_mm_update_gaussian_transform_with_pod", i32, vp, P(GaussianTransformPod))
    sig("sb_mm_set_selection", i32, vp, u64, vp, vp, u64, i32)
    sig("sb_mm_render", i32, vp, vp, P(Target), vp, u32)
    sig("sb_mm_read_model_indices", i32, vp, u64, vp, vp, u64, P(DrawIndirectArgs))
    sig("sb_mm_update_camera", i32, vp, vp, f32, f32, f32, f32, f32, u32, u32)
    sig("sb_mm_update_model_transform", i32, vp, u64, vp, vp, vp)
    sig("sb_mm_update_gaussian_transform", i32, vp, f32, i32, i32, i32, f32)
    sig("sb_mm_insert_model_from_gaussians", i32, vp, u64, vp, u64, P(i32))
    sig("sb_mm_insert_model_from_device", i32, vp, u64, vp, u64, u64, P(i32))
    sig("sb_mm_select_rect", i32, vp, u64, vp, f32, f32, f32, f32)
    sig("sb_mm_select_brush", i32, vp, u64, vp, P(f32), u32, f32, i32)
    sig("sb_mm_enable_selection", i32, vp, u64, i32, i32)
    sig("sb_mm_read_selection", i32, vp, u64, vp, vp, u64)
    sig("sb_mm_render_with_pass", i32, vp, vp, P(Target), P(DepthAttachment), i32, vp, u32)
    _lib = l
    return l


def _check(status: int, ctx=None):
    if status != 0:
        msg = load().sb_last_error_string(ctx).decode()
        raise SplatError(status, msg or load().sb_status_string(status).decode())


def _f(v):
    return np.ascontiguousarray(v, dtype=np.float32)


# ---------------------------------------------------------------- host helpers (no GPU)

def pod_stride(sh_fmt: int, cov_fmt: int) -> int:
    return load().sb_pod_stride(sh_fmt, cov_fmt)


def padded_key_count(n: int) -> int:
    return load().sb_padded_key_count(n)


def keys_buffer_size_bytes(n: int) -> int:
    return load().sb_keys_buffer_size_bytes(n)


def pack_gaussians(gaussians: np.ndarray, sh_fmt: int = SH_SINGLE, cov_fmt: int = COV_SINGLE) -> np.ndarray:
    g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN_DTYPE)
    out = np.zeros(len(g) * pod_stride(sh_fmt, cov_fmt), dtype=np.uint8)
    _check(load().sb_pack_gaussians(g.ctypes.data, len(g), sh_fmt, cov_fmt, out.ctypes.data))
    return out


def camera_pod(pos, yaw, pitch, width, height, z_near=0.1, z_far=1e4, fov_y=float(np.deg2rad(np.float32(60.0)))) -> CameraPod:
    pod = CameraPod()
    p = _f(pos)
    _check(load().sb_camera_pod(p.ctypes.data, yaw, pitch, z_near, z_far, fov_y, width, height, C.byref(pod)))
    return pod


def model_transform_pod(pos=(0, 0, 0), rot=(0, 0, 0, 1), scale=(1, 1, 1)) -> ModelTransformPod:
    pod = ModelTransformPod()
    a, b, c = _f(pos), _f(rot), _f(scale)
    _check(load().sb_model_transform_pod(a.ctypes.data, b.ctypes.data, c.ctypes.data, C.byref(pod)))
    return pod


def gaussian_transform_pod(size=1.0, display_mode=MODE_SPLAT, sh_deg=3, no_sh0=False, max_std_dev=3.0) -> GaussianTransformPod:
    pod = GaussianTransformPod()
    _check(load().sb_gaussian_transform_pod(size, display_mode, sh_deg, int(no_sh0), max_std_dev, C.byref(pod)))
    return pod


def read_ply(path: str) -> np.ndarray:
    ptr, n = C.c_void_p(), C.c_uint64()
    _check(load().sb_read_ply(path.encode(), C.byref(ptr), C.byref(n)))
    try:
        buf = (C.c_uint8 * (n.value * GAUSSIAN_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=GAUSSIAN_DTYPE).copy()
    finally:
        load().sb_free(ptr)


def read_spz(path: str) -> np.ndarray:
    """`Gaussians::read_from_file(path, GaussiansSource::Spz)`: Niantic .spz v2 / v3 -> Gaussian records."""
    ptr, n = C.c_void_p(), C.c_uint64()
    _check(load().sb_read_spz(path.encode(), C.byref(ptr), C.byref(n)))
    try:
        buf = (C.c_uint8 * (n.value * GAUSSIAN_DTYPE.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=GAUSSIAN_DTYPE).copy()
    finally:
        load().sb_free(ptr)


# ---------------------------------------------------------------- device objects

_BPP = {TARGET_RGBA8: 4, TARGET_BGRA8: 4, TARGET_RGBA16F: 8, TARGET_RGBA32F: 16, TARGET_RGBA8_SRGB: 4, TARGET_BGRA8_SRGB: 4}


def _stream_handle(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)  # torch.cuda.Stream


COMPARE_NEVER, COMPARE_LESS, COMPARE_EQUAL, COMPARE_LESS_EQUAL, COMPARE_GREATER, COMPARE_NOT_EQUAL, COMPARE_GREATER_EQUAL, COMPARE_ALWAYS = range(1, 9)


def make_target(tensor_or_ptr, width, height, fmt, pitch=None, row0=0, rows=0) -> Target:
    ptr = tensor_or_ptr if isinstance(tensor_or_ptr, int) else tensor_or_ptr.data_ptr()
    t = Target()
    t.d_pixels = ptr
    t.pitch_bytes = pitch if pitch is not None else width * _BPP[fmt]
    t.width, t.height, t.format, t.row0, t.rows = width, height, fmt, row0, rows
    return t


class Context:
    """Replaces the `&wgpu::Device` every reference constructor takes."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(load().sb_ctx_create(device, C.byref(self._h)))
        self.device = device

    def probe_peaks(self, stream=None):
        """(FP32 lane-ops/s with FMA = 1, shared-memory bytes/s) measured on this device by two microbenchmarks."""
        a, b = C.c_double(), C.c_double()
        _check(load().sb_probe_peaks(self._h, _stream_handle(stream), C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def set_model_size_limit(self, nbytes: int):
        _check(load().sb_ctx_set_model_size_limit(self._h, nbytes), self._h)

    def close(self):
        if self._h:
            load().sb_ctx_destroy(self._h)
            self._h = C.c_void_p()


class SharedFrame:
    """A device frame buffer the other GPUs of the node render their strips into (sb_shared_frame_*, CUDA IPC over NVLink).
    The owner creates it and publishes `handle` (64 bytes); peers construct it from that handle.  `ptr` is a device pointer
    valid on the calling process's GPU."""

    def __init__(self, ctx: Context, nbytes: int = 0, handle: bytes | None = None):
        self.ctx, self.nbytes, self.owner = ctx, nbytes, handle is None
        p = C.c_void_p()
        if handle is None:
            buf = C.create_string_buffer(64)
            _check(load().sb_shared_frame_create(ctx._h, nbytes, C.byref(p), buf), ctx._h)
            self.handle = buf.raw
        else:
            self.handle = bytes(handle)
            _check(load().sb_shared_frame_open(ctx._h, self.handle, C.byref(p)), ctx._h)
        self.ptr = int(p.value)

    def close(self):
        if self.ptr:
            fn = load().sb_shared_frame_destroy if self.owner else load().sb_shared_frame_close
            _check(fn(self.ctx._h, self.ptr), self.ctx._h)
            self.ptr = 0


class Strips:
    """sb_strips_*: one frame as screen strips with the Preprocessor's work partitioned over the ranks as well.  Construct on every
    rank, exchange `exported` (192 bytes) between all ranks, `connect(all_exports_in_rank_order)`; per frame `scatter(stream)`,
    a barrier across the ranks, `render(stream, target...)`, and a second barrier before the next frame's scatter."""

    EXPORT_BYTES = 192

    def __init__(self, viewer: "Viewer", world: int, rank: int, bounds):
        self.viewer, self.world, self.rank = viewer, world, rank
        r0 = (C.c_uint32 * world)(*[int(b[0]) for b in bounds])
        rn = (C.c_uint32 * world)(*[int(b[1]) for b in bounds])
        self._h = C.c_void_p()
        buf = C.create_string_buffer(self.EXPORT_BYTES)
        _check(load().sb_strips_create(viewer._h, world, rank, r0, rn, C.byref(self._h), buf), viewer.ctx._h)
        self.exported = buf.raw

    def connect(self, all_exports):
        blob = b"".join(bytes(e) for e in all_exports)
        assert len(blob) == self.EXPORT_BYTES * self.world
        _check(load().sb_strips_connect(self._h, blob), self.viewer.ctx._h)

    def scatter(self, stream=None):
        _check(load().sb_strips_scatter(self._h, _stream_handle(stream)), self.viewer.ctx._h)

    def render(self, target, width, height, row0, rows, pitch=None, stream=None):
        t = make_target(target, width, height, self.viewer.target_format, pitch, row0, rows)
        _check(load().sb_strips_render(self._h, _stream_handle(stream), C.byref(t)), self.viewer.ctx._h)

    def close(self):
        if self._h:
            load().sb_strips_destroy(self._h)
            self._h = C.c_void_p()


class Viewer:
    """`wgpu_3dgs_viewer::Viewer<G>` (reference src/lib.rs:65-276)."""

    def __init__(self, ctx: Context, pods: np.ndarray | None = None, n: int | None = None, *, gaussians: np.ndarray | None = None,
                 sh_fmt=SH_SINGLE, cov_fmt=COV_SINGLE, target_format=TARGET_RGBA8, device_pods=None):
        self.ctx, self.sh_fmt, self.cov_fmt, self.target_format = ctx, sh_fmt, cov_fmt, target_format
        self._h = C.c_void_p()
        l = load()
        if gaussians is not None:
            g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN_DTYPE)
            self.n = len(g)
            _check(l.sb_viewer_create_from_gaussians(ctx._h, sh_fmt, cov_fmt, target_format, g.ctypes.data, self.n, C.byref(self._h)), ctx._h)
        elif device_pods is not None:  # torch uint8 tensor on the device, adopted (not copied)
            self.n = int(n)
            self._keepalive = device_pods
            _check(l.sb_viewer_create_from_device(ctx._h, sh_fmt, cov_fmt, target_format, device_pods.data_ptr(),
                                                  device_pods.numel() * device_pods.element_size(), self.n, C.byref(self._h)), ctx._h)
        else:
            p = np.ascontiguousarray(pods, dtype=np.uint8)
            self.n = int(n if n is not None else p.size // pod_stride(sh_fmt, cov_fmt))
            if p.size != self.n * pod_stride(sh_fmt, cov_fmt):
                raise SplatError(5, "pods size != n * stride")
            _check(l.sb_viewer_create(ctx._h, sh_fmt, cov_fmt, target_format, p.ctypes.data, self.n, C.byref(self._h)), ctx._h)

    def close(self):
        if self._h:
            load().sb_viewer_destroy(self._h)
            self._h = C.c_void_p()

    # --- updates (src/lib.rs:202-263)
    def update_camera_with_pod(self, pod: CameraPod):
        _check(load().sb_viewer_update_camera_with_pod(self._h, C.byref(pod)), self.ctx._h)

    def update_camera(self, pos, yaw, pitch, width, height, z_near=0.1, z_far=1e4, fov_y=float(np.deg2rad(np.float32(60.0)))):
        p = _f(pos)
        _check(load().sb_viewer_update_camera(self._h, p.ctypes.data, yaw, pitch, z_near, z_far, fov_y, width, height), self.ctx._h)

    def update_model_transform(self, pos=(0, 0, 0), rot=(0, 0, 0, 1), scale=(1, 1, 1)):
        a, b, c = _f(pos), _f(rot), _f(scale)
        _check(load().sb_viewer_update_model_transform(self._h, a.ctypes.data, b.ctypes.data, c.ctypes.data), self.ctx._h)

    def update_model_transform_with_pod(self, pod: ModelTransformPod):
        _check(load().sb_viewer_update_model_transform_with_pod(self._h, C.byref(pod)), self.ctx._h)

    def update_gaussian_transform(self, size=1.0, display_mode=MODE_SPLAT, sh_deg=3, no_sh0=False, max_std_dev=3.0):
        _check(load().sb_viewer_update_gaussian_transform(self._h, size, display_mode, sh_deg, int(no_sh0), max_std_dev), self.ctx._h)

    def update_gaussian_transform_with_pod(self, pod: GaussianTransformPod):
        _check(load().sb_viewer_update_gaussian_transform_with_pod(self._h, C.byref(pod)), self.ctx._h)

    # --- selection mask (src/lib.rs:134-144, src/selection/buffer.rs:147-172)
    def enable_selection(self, enabled=True):
        _check(load().sb_viewer_enable_selection(self._h, int(enabled)), self.ctx._h)

    def set_selection(self, words: np.ndarray, stream=None):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        _check(load().sb_viewer_set_selection(self._h, _stream_handle(stream), w.ctypes.data, len(w)), self.ctx._h)

    def set_invert_selection(self, invert: bool):
        _check(load().sb_viewer_set_invert_selection(self._h, int(invert)), self.ctx._h)

    def select_rect(self, x0, y0, x1, y1, stream=None):
        _check(load().sb_viewer_select_rect(self._h, _stream_handle(stream), x0, y0, x1, y1), self.ctx._h)

    def select_brush(self, points, radius, accumulate=False, stream=None):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        _check(load().sb_viewer_select_brush(self._h, _stream_handle(stream), pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts),
                                             float(radius), int(accumulate)), self.ctx._h)

    def apply_rgb_override(self, rgb, alpha=1.0, stream=None):
        """editor NonDestructiveModifier + rgb override on the selected Gaussians (tests/e2e/selection.rs:54-116)."""
        c = _f(rgb)
        _check(load().sb_viewer_apply_rgb_override(self._h, _stream_handle(stream), c.ctypes.data_as(C.POINTER(C.c_float)), float(alpha)), self.ctx._h)

    def apply_basic_color_modifiers(self, rgb=None, hsv=(0.0, 1.0, 1.0), alpha=1.0, contrast=0.0, exposure=0.0, gamma=1.0, stream=None):
        """editor BasicColorModifiers on the selected Gaussians: rgb override (rgb=...) or HSV (hue shift in turns, saturation
        scale, value scale), then contrast, exposure (stops), gamma and the alpha scale."""
        m = BasicColorModifiers()
        m.rgb_override = int(rgb is not None)
        for i, x in enumerate(rgb if rgb is not None else hsv):
            m.rgb_or_hsv[i] = float(x)
        m.alpha, m.contrast, m.exposure, m.gamma = float(alpha), float(contrast), float(exposure), float(gamma)
        _check(load().sb_viewer_apply_basic_color_modifiers(self._h, _stream_handle(stream), C.byref(m)), self.ctx._h)

    def restore_gaussians(self, stream=None):
        _check(load().sb_viewer_restore_gaussians(self._h, _stream_handle(stream)), self.ctx._h)

    def read_selection(self, stream=None) -> np.ndarray:
        out = np.zeros((self.n + 31) // 32, dtype=np.uint32)
        _check(load().sb_viewer_read_selection(self._h, _stream_handle(stream), out.ctypes.data, len(out)), self.ctx._h)
        return out

    # --- the frame (src/lib.rs:266-275) and its three stages
    def render(self, target, width, height, stream=None, row0=0, rows=0, pitch=None):
        t = make_target(target, width, height, self.target_format, pitch, row0, rows)
        _check(load().sb_viewer_render(self._h, _stream_handle(stream), C.byref(t)), self.ctx._h)

    def render_with_pass(self, target, width, height, depth=None, compare=COMPARE_ALWAYS, depth_write=False, load_target=True,
                         run_stages=True, stream=None):
        """Renderer::render_with_pass + depth_stencil: composite over the target's contents against an f32 depth tensor."""
        t = make_target(target, width, height, self.target_format)
        d = None
        if depth is not None:
            d = DepthAttachment()
            d.d_depth = depth.data_ptr()
            d.pitch_bytes = width * 4
            d.compare = compare
            d.write_enabled = int(depth_write)
        _check(load().sb_viewer_render_with_pass(self._h, _stream_handle(stream), C.byref(t), C.byref(d) if d is not None else None,
                                                 int(load_target), int(run_stages)), self.ctx._h)

    def preprocess(self, stream=None):
        _check(load().sb_viewer_preprocess(self._h, _stream_handle(stream)), self.ctx._h)

    def sort(self, stream=None):
        _check(load().sb_viewer_sort(self._h, _stream_handle(stream)), self.ctx._h)

    def draw(self, target, width, height, stream=None, row0=0, rows=0, pitch=None):
        t = make_target(target, width, height, self.target_format, pitch, row0, rows)
        _check(load().sb_viewer_draw(self._h, _stream_handle(stream), C.byref(t)), self.ctx._h)

    def render_to_host(self, cam: CameraPod, host_ptr: int, host_bytes: int, stream=None):
        _check(load().sb_viewer_render_to_host(self._h, _stream_handle(stream), C.byref(cam), host_ptr, host_bytes), self.ctx._h)

    def render_batch(self, cams, targets=None, width=None, height=None, host_ptrs=None, stream=None):
        """Render len(cams) views (pipelined two deep).  targets: device tensors/pointers, or host_ptrs: pinned host pointers."""
        n = len(cams)
        cam_arr = (CameraPod * n)(*cams)
        if targets is not None:
            tg = (Target * n)(*[make_target(t, width, height, self.target_format) for t in targets])
            _check(load().sb_viewer_render_batch(self._h, _stream_handle(stream), cam_arr, tg, None, n), self.ctx._h)
        else:
            hp = (C.c_void_p * n)(*host_ptrs)
            _check(load().sb_viewer_render_batch(self._h, _stream_handle(stream), cam_arr, None, hp, n), self.ctx._h)

    # --- artefact read-back (synchronising)
    def read_indirect_args(self, stream=None):
        d, s = DrawIndirectArgs(), DispatchIndirectArgs()
        _check(load().sb_viewer_read_indirect_args(self._h, _stream_handle(stream), C.byref(d), C.byref(s)), self.ctx._h)
        return (np.array([d.vertex_count, d.instance_count, d.first_vertex, d.first_instance], dtype=np.uint32),
                np.array([s.x, s.y, s.z], dtype=np.uint32))

    def read_indices(self, count: int, stream=None) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint32)
        _check(load().sb_viewer_read_indices(self._h, _stream_handle(stream), out.ctypes.data, count), self.ctx._h)
        return out

    def read_depth_keys(self, count: int, stream=None) -> np.ndarray:
        out = np.zeros(count, dtype=np.float32)
        _check(load().sb_viewer_read_depth_keys(self._h, _stream_handle(stream), out.ctypes.data, count), self.ctx._h)
        return out

    def read_frame_stats(self, stream=None):
        v, d, o = C.c_uint64(), C.c_uint64(), C.c_uint32()
        st = load().sb_viewer_read_frame_stats(self._h, _stream_handle(stream), C.byref(v), C.byref(d), C.byref(o))
        if st not in (0, 7):
            _check(st, self.ctx._h)
        return dict(visible=v.value, duplicates=d.value, overflowed=bool(o.value))

    def device_pointers(self):
        l, out = load(), {}
        p, b = C.c_void_p(), C.c_uint64()
        _check(l.sb_viewer_gaussians_ptr(self._h, C.byref(p), C.byref(b)), self.ctx._h)
        out["gaussians"] = (p.value, b.value)
        _check(l.sb_viewer_indirect_args_ptr(self._h, C.byref(p)), self.ctx._h)
        out["indirect_args"] = (p.value, 16)
        _check(l.sb_viewer_radix_sort_indirect_args_ptr(self._h, C.byref(p)), self.ctx._h)
        out["radix_sort_indirect_args"] = (p.value, 12)
        _check(l.sb_viewer_indirect_indices_ptr(self._h, C.byref(p), C.byref(b)), self.ctx._h)
        out["indirect_indices"] = (p.value, b.value * 4)
        _check(l.sb_viewer_gaussians_depth_ptr(self._h, C.byref(p), C.byref(b)), self.ctx._h)
        out["gaussians_depth"] = (p.value, b.value)
        return out

    def raster_path(self) -> str:
        f = C.c_int32()
        _check(load().sb_viewer_raster_path(self._h, C.byref(f)), self.ctx._h)
        return "tma_gather4" if f.value else "bulk"

    def set_strict_exp(self, strict: bool):
        _check(load().sb_viewer_set_strict_exp(self._h, int(strict)), self.ctx._h)

    def set_exact_cutoff(self, enabled: bool):
        _check(load().sb_viewer_set_exact_cutoff(self._h, int(enabled)), self.ctx._h)

    def read_tile_row_work(self, n_rows: int, stream=None) -> np.ndarray:
        """(splat, tile) duplicates per 16-pixel tile row of the last binned (full) frame."""
        out = np.zeros(n_rows, dtype=np.uint64)
        _check(load().sb_viewer_read_tile_row_work(self._h, _stream_handle(stream), out.ctypes.data_as(C.POINTER(C.c_uint64)), n_rows), self.ctx._h)
        return out

    def set_strip_cull(self, enabled: bool):
        """Strip renders keep only the visible splats whose tile box meets the strip (one frame sharded over GPUs)."""
        _check(load().sb_viewer_set_strip_cull(self._h, int(enabled)), self.ctx._h)

    STAGES = ("preprocess", "depth_sort", "tile_emit", "tile_sort", "gather", "raster")

    def set_raster_counting(self, enabled: bool):
        _check(load().sb_viewer_set_raster_counting(self._h, int(enabled)), self.ctx._h)

    def read_raster_counters(self, stream=None) -> dict:
        a, e = C.c_uint64(), C.c_uint64()
        _check(load().sb_viewer_read_raster_counters(self._h, _stream_handle(stream), C.byref(a), C.byref(e)), self.ctx._h)
        w, wa = C.c_uint64(), C.c_uint64()
        _check(load().sb_viewer_read_raster_warp_counters(self._h, _stream_handle(stream), C.byref(w), C.byref(wa)), self.ctx._h)
        return dict(alive=a.value, evaluated=e.value, warp_evals=w.value, warp_evals_alive=wa.value)

    def set_stage_timing(self, enabled: bool):
        _check(load().sb_viewer_set_stage_timing(self._h, int(enabled)), self.ctx._h)

    def read_stage_times(self, stream=None) -> dict:
        ms = (C.c_float * 6)()
        _check(load().sb_viewer_read_stage_times(self._h, _stream_handle(stream), ms), self.ctx._h)
        return dict(zip(self.STAGES, [float(x) for x in ms]))

    def reserve_duplicates(self, capacity: int):
        _check(load().sb_viewer_reserve_duplicates(self._h, capacity), self.ctx._h)


class RadixSorter:
    """`RadixSorter<()>` (reference src/radix_sorter.rs:71-96): stable (u32 key, u32 payload) sort."""

    def __init__(self, ctx: Context, capacity: int):
        self.ctx, self.capacity = ctx, capacity
        self._h = C.c_void_p()
        _check(load().sb_sorter_create(ctx._h, capacity, C.byref(self._h)), ctx._h)

    def sort(self, d_keys: int, d_payload: int, d_count: int, max_count: int, begin_bit=0, end_bit=32, stream=None):
        _check(load().sb_sorter_sort(self._h, _stream_handle(stream), d_keys, d_payload, d_count, max_count, begin_bit, end_bit), self.ctx._h)

    def close(self):
        if self._h:
            load().sb_sorter_destroy(self._h)
            self._h = C.c_void_p()


def _nbytes(t) -> int:
    return t.numel() * t.element_size()


class Preprocessor:
    """`Preprocessor<G, ()>` (reference src/preprocessor.rs:370-450): the stage alone, on the caller's device buffers
    (torch tensors here)."""

    def __init__(self, ctx: Context, n: int, sh_fmt=SH_SINGLE, cov_fmt=COV_SINGLE):
        self.ctx, self.n = ctx, n
        self._h = C.c_void_p()
        _check(load().sb_preprocessor_create(ctx._h, sh_fmt, cov_fmt, n, C.byref(self._h)), ctx._h)

    @staticmethod
    def create_bind_group(camera, model_transform, gaussian_transform, gaussians, indirect_args, radix_sort_indirect_args,
                          indirect_indices, gaussians_depth, selection=None, invert_selection=1) -> PreprocessorBindGroup:
        bg = PreprocessorBindGroup()
        bg.camera, bg.model_transform, bg.gaussian_transform = camera, model_transform, gaussian_transform
        bg.d_gaussians, bg.gaussians_bytes = gaussians.data_ptr(), _nbytes(gaussians)
        bg.d_indirect_args = indirect_args.data_ptr()
        bg.d_radix_sort_indirect_args = radix_sort_indirect_args.data_ptr()
        bg.d_indirect_indices, bg.indirect_indices_bytes = indirect_indices.data_ptr(), _nbytes(indirect_indices)
        bg.d_gaussians_depth, bg.gaussians_depth_bytes = gaussians_depth.data_ptr(), _nbytes(gaussians_depth)
        bg.d_selection = selection.data_ptr() if selection is not None else None
        bg.invert_selection = int(invert_selection)
        bg._keepalive = (gaussians, indirect_args, radix_sort_indirect_args, indirect_indices, gaussians_depth, selection)
        return bg

    def preprocess(self, bind_group: PreprocessorBindGroup, gaussian_count: int | None = None, stream=None):
        _check(load().sb_preprocessor_preprocess(self._h, _stream_handle(stream), C.byref(bind_group),
                                                 self.n if gaussian_count is None else gaussian_count), self.ctx._h)

    def close(self):
        if self._h:
            load().sb_preprocessor_destroy(self._h)
            self._h = C.c_void_p()


class Renderer:
    """`Renderer<G, ()>` (reference src/renderer.rs:242-356): one indirect instanced draw of the caller's indices."""

    def __init__(self, ctx: Context, n: int, sh_fmt=SH_SINGLE, cov_fmt=COV_SINGLE, target_format=TARGET_RGBA8):
        self.ctx, self.n, self.target_format = ctx, n, target_format
        self._h = C.c_void_p()
        _check(load().sb_renderer_create(ctx._h, sh_fmt, cov_fmt, target_format, n, C.byref(self._h)), ctx._h)

    @staticmethod
    def create_bind_group(camera, model_transform, gaussian_transform, gaussians, indirect_indices) -> RendererBindGroup:
        bg = RendererBindGroup()
        bg.camera, bg.model_transform, bg.gaussian_transform = camera, model_transform, gaussian_transform
        bg.d_gaussians, bg.gaussians_bytes = gaussians.data_ptr(), _nbytes(gaussians)
        bg.d_indirect_indices, bg.indirect_indices_bytes = indirect_indices.data_ptr(), _nbytes(indirect_indices)
        bg._keepalive = (gaussians, indirect_indices)
        return bg

    def set_strict_exp(self, strict: bool):
        _check(load().sb_renderer_set_strict_exp(self._h, int(strict)), self.ctx._h)

    def render(self, target, width, height, bind_group: RendererBindGroup, indirect_args, depth=None, compare=COMPARE_ALWAYS,
               depth_write=False, load_target=False, stream=None):
        """render (load_target False: clear to BLACK) / render_with_pass (load_target True) with `indirect_args` a device
        tensor holding DrawIndirectArgs."""
        t = make_target(target, width, height, self.target_format)
        d = None
        if depth is not None:
            d = DepthAttachment()
            d.d_depth, d.pitch_bytes, d.compare, d.write_enabled = depth.data_ptr(), width * 4, compare, int(depth_write)
        _check(load().sb_renderer_render(self._h, _stream_handle(stream), C.byref(bind_group), C.byref(t), indirect_args.data_ptr(),
                                         C.byref(d) if d is not None else None, int(load_target)), self.ctx._h)

    def close(self):
        if self._h:
            load().sb_renderer_destroy(self._h)
            self._h = C.c_void_p()


class MultiModelViewer:
    """`MultiModelViewer<G, K = u64>` (reference src/multi_model.rs:291-531)."""

    def __init__(self, ctx: Context, sh_fmt=SH_SINGLE, cov_fmt=COV_SINGLE, target_format=TARGET_RGBA8):
        self.ctx, self.sh_fmt, self.cov_fmt, self.target_format = ctx, sh_fmt, cov_fmt, target_format
        self._h = C.c_void_p()
        _check(load().sb_mm_create(ctx._h, sh_fmt, cov_fmt, target_format, C.byref(self._h)), ctx._h)

    def insert_model(self, key: int, pods: np.ndarray, n: int) -> bool:
        p = np.ascontiguousarray(pods, dtype=np.uint8)
        rep = C.c_int32()
        _check(load().sb_mm_insert_model(self._h, key, p.ctypes.data, n, C.byref(rep)), self.ctx._h)
        return bool(rep.value)

    def remove_model(self, key: int) -> bool:
        rem = C.c_int32()
        _check(load().sb_mm_remove_model(self._h, key, C.byref(rem)), self.ctx._h)
        return bool(rem.value)

    def update_camera_with_pod(self, pod: CameraPod):
        _check(load().sb_mm_update_camera_with_pod(self._h, C.byref(pod)), self.ctx._h)

    def update_model_transform_with_pod(self, key: int, pod: ModelTransformPod):
        _check(load().sb_mm_update_model_transform_with_pod(self._h, key, C.byref(pod)), self.ctx._h)

    def update_gaussian_transform_with_pod(self, pod: GaussianTransformPod):
        _check(load().sb_mm_update_gaussian_transform_with_pod(self._h, C.byref(pod)), self.ctx._h)

    def set_selection(self, key: int, words: np.ndarray | None, invert: bool = True, stream=None):
        if words is None:
            _check(load().sb_mm_set_selection(self._h, key, _stream_handle(stream), None, 0, int(invert)), self.ctx._h)
            return
        w = np.ascontiguousarray(words, dtype=np.uint32)
        _check(load().sb_mm_set_selection(self._h, key, _stream_handle(stream), w.ctypes.data, len(w), int(invert)), self.ctx._h)

    def render(self, target, width, height, keys, stream=None):
        t = make_target(target, width, height, self.target_format)
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        _check(load().sb_mm_render(self._h, _stream_handle(stream), C.byref(t), k.ctypes.data, len(k)), self.ctx._h)

    def read_model_indices(self, key: int, count: int, stream=None):
        out = np.zeros(count, dtype=np.uint32)
        d = DrawIndirectArgs()
        _check(load().sb_mm_read_model_indices(self._h, key, _stream_handle(stream), out.ctypes.data, count, C.byref(d)), self.ctx._h)
        return out, d.instance_count

    # --- non-pod updates (src/multi_model.rs:398-464)
    def update_camera(self, pos, yaw, pitch, width, height, z_near=0.1, z_far=1e4, fov_y=float(np.deg2rad(np.float32(60.0)))):
        p = _f(pos)
        _check(load().sb_mm_update_camera(self._h, p.ctypes.data, yaw, pitch, z_near, z_far, fov_y, width, height), self.ctx._h)

    def update_model_transform(self, key: int, pos=(0, 0, 0), rot=(0, 0, 0, 1), scale=(1, 1, 1)):
        a, b, c = _f(pos), _f(rot), _f(scale)
        _check(load().sb_mm_update_model_transform(self._h, key, a.ctypes.data, b.ctypes.data, c.ctypes.data), self.ctx._h)

    def update_gaussian_transform(self, size=1.0, display_mode=MODE_SPLAT, sh_deg=3, no_sh0=False, max_std_dev=3.0):
        _check(load().sb_mm_update_gaussian_transform(self._h, size, display_mode, sh_deg, int(no_sh0), max_std_dev), self.ctx._h)

    def insert_model_from_gaussians(self, key: int, gaussians: np.ndarray) -> bool:
        g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN_DTYPE)
        rep = C.c_int32()
        _check(load().sb_mm_insert_model_from_gaussians(self._h, key, g.ctypes.data, len(g), C.byref(rep)), self.ctx._h)
        return bool(rep.value)

    def insert_model_from_device(self, key: int, device_pods, n: int) -> bool:
        rep = C.c_int32()
        self._keepalive = getattr(self, "_keepalive", {})
        self._keepalive[key] = device_pods
        _check(load().sb_mm_insert_model_from_device(self._h, key, device_pods.data_ptr(), _nbytes(device_pods), n, C.byref(rep)), self.ctx._h)
        return bool(rep.value)

    # --- per-model selection buffers (src/multi_model.rs:81-84)
    def select_rect(self, key: int, x0, y0, x1, y1, stream=None):
        _check(load().sb_mm_select_rect(self._h, key, _stream_handle(stream), x0, y0, x1, y1), self.ctx._h)

    def select_brush(self, key: int, points, radius, accumulate=False, stream=None):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 2)
        _check(load().sb_mm_select_brush(self._h, key, _stream_handle(stream), pts.ctypes.data_as(C.POINTER(C.c_float)), len(pts),
                                         float(radius), int(accumulate)), self.ctx._h)

    def enable_selection(self, key: int, enabled=True, invert=True):
        _check(load().sb_mm_enable_selection(self._h, key, int(enabled), int(invert)), self.ctx._h)

    def read_selection(self, key: int, n: int, stream=None) -> np.ndarray:
        out = np.zeros((n + 31) // 32, dtype=np.uint32)
        _check(load().sb_mm_read_selection(self._h, key, _stream_handle(stream), out.ctypes.data, len(out)), self.ctx._h)
        return out

    def render_with_pass(self, target, width, height, keys, depth=None, compare=COMPARE_ALWAYS, depth_write=False, load_target=True,
                         stream=None):
        """new_with_options(depth_stencil) + renderer.render_with_pass per model inside a caller's pass."""
        t = make_target(target, width, height, self.target_format)
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        d = None
        if depth is not None:
            d = DepthAttachment()
            d.d_depth, d.pitch_bytes, d.compare, d.write_enabled = depth.data_ptr(), width * 4, compare, int(depth_write)
        _check(load().sb_mm_render_with_pass(self._h, _stream_handle(stream), C.byref(t), C.byref(d) if d is not None else None,
                                             int(load_target), k.ctypes.data, len(k)), self.ctx._h)

    def close(self):
        if self._h:
            load().sb_mm_destroy(self._h)
            self._h = C.c_void_p()
