"""splat_b200 — Python harness over the C ABI (include/splat_b200.h).

The product is the C-ABI shared library `lib/libsplat_b200.so` (hand-written sm_100a CUDA).
This package is the thin ctypes mirror of the reference's public API for the hot path
(`Viewer::new / update_camera / render`, stage access, `MultiModelViewer`, selection mask) that
tests/ and bench.py drive.  There is NO CPU fallback: if the library is missing, importing
`load()` raises; if no B200 is present, `Context()` raises.
"""
from .api import (  # noqa: F401
    COV_HALF, COV_ROT_SCALE, COV_SINGLE, MODE_ELLIPSE, MODE_POINT, MODE_SPLAT, SH_HALF, SH_NONE, SH_NORM8,
    SH_SINGLE, TARGET_BGRA8, TARGET_RGBA8, TARGET_RGBA16F, TARGET_RGBA32F, TARGET_RGBA8_SRGB, TARGET_BGRA8_SRGB, GAUSSIAN_DTYPE, CameraPod,
    Context, GaussianTransformPod, ModelTransformPod, MultiModelViewer, RadixSorter, SplatError, Viewer,
    build, camera_pod, gaussian_transform_pod, lib_path, load, model_transform_pod, pack_gaussians,
    pod_stride, read_ply, read_spz, padded_key_count, keys_buffer_size_bytes, EXPORTED_SYMBOLS, DepthAttachment,
    COMPARE_NEVER, COMPARE_LESS, COMPARE_EQUAL, COMPARE_LESS_EQUAL, COMPARE_GREATER, COMPARE_NOT_EQUAL, COMPARE_GREATER_EQUAL,
    COMPARE_ALWAYS, Preprocessor, Renderer, PreprocessorBindGroup, RendererBindGroup, DrawIndirectArgs, DispatchIndirectArgs,
)
from . import scenes, sharding  # noqa: F401
