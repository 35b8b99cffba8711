//! `splat-b200`: the per-frame splat pipeline of `wgpu-3dgs-viewer` on an NVIDIA B200, behind the
//! same shape of API (`Viewer::new / update_camera / render`, stage access, `MultiModelViewer`).
//!
//! Where the reference takes `&wgpu::Device`, `&wgpu::Queue`, `&mut wgpu::CommandEncoder` and a
//! `&wgpu::TextureView`, this crate takes a [`Context`], nothing, a CUDA stream and a [`Target`]
//! (caller-owned device memory).  `render` only enqueues work, like recording into an encoder.
//!
//! SOURCE ONLY — never compiled in the build container (no Rust toolchain there).
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_void, CStr};
use std::marker::PhantomData;

pub mod ffi {
    use super::*;
    #[repr(C)] pub struct SbContext { _p: [u8; 0] }
    #[repr(C)] pub struct SbViewer { _p: [u8; 0] }
    #[repr(C)] pub struct SbMultiModelViewer { _p: [u8; 0] }
    #[repr(C)] pub struct SbPreprocessor { _p: [u8; 0] }
    #[repr(C)] pub struct SbRadixSorter { _p: [u8; 0] }
    #[repr(C)] pub struct SbRenderer { _p: [u8; 0] }

    /// `CameraPod` — byte-identical to the reference (src/buffer/camera.rs:63-80).
    #[repr(C)] #[derive(Clone, Copy, Debug, PartialEq, bytemuck::Pod, bytemuck::Zeroable)]
    pub struct CameraPod { pub view: [f32; 16], pub proj: [f32; 16], pub size: [f32; 2], pub _padding: [u32; 2] }
    #[repr(C)] #[derive(Clone, Copy, Debug, PartialEq, bytemuck::Pod, bytemuck::Zeroable)]
    pub struct ModelTransformPod { pub pos: [f32; 3], pub _pad0: f32, pub rot: [f32; 4], pub scale: [f32; 3], pub _pad1: f32 }
    #[repr(C)] #[derive(Clone, Copy, Debug, PartialEq, bytemuck::Pod, bytemuck::Zeroable)]
    pub struct GaussianTransformPod { pub size: f32, pub display_mode: u8, pub sh_deg: u8, pub no_sh0: u8, pub max_std_dev: u8 }
    #[repr(C)] #[derive(Clone, Copy, Debug)]
    pub struct Gaussian { pub pos: [f32; 3], pub color: [u8; 4], pub sh: [f32; 45], pub scale: [f32; 3], pub rot: [f32; 4] }
    #[repr(C)] #[derive(Clone, Copy, Debug, Default)]
    pub struct DrawIndirectArgs { pub vertex_count: u32, pub instance_count: u32, pub first_vertex: u32, pub first_instance: u32 }
    #[repr(C)] #[derive(Clone, Copy, Debug, Default)]
    pub struct DispatchIndirectArgs { pub x: u32, pub y: u32, pub z: u32 }
    #[repr(C)] #[derive(Clone, Copy, Debug)]
    pub struct Target { pub d_pixels: *mut c_void, pub pitch_bytes: u32, pub width: u32, pub height: u32, pub format: i32, pub row0: u32, pub rows: u32 }

    /// `wgpu::DepthStencilState` of `ViewerCreateOptions::depth_stencil` reduced to what the draw needs, plus the
    /// pass's Depth32Float attachment (reference src/lib.rs:279-284, src/renderer.rs:123,304).
    #[repr(C)] #[derive(Clone, Copy, Debug)]
    pub struct DepthAttachment { pub d_depth: *mut c_void, pub pitch_bytes: u32, pub compare: i32, pub write_enabled: i32 }

    /// The `Preprocessor` bind group (reference src/preprocessor.rs:104-221): bindings 0-9 on caller-owned device buffers.
    #[repr(C)] #[derive(Clone, Copy)]
    pub struct PreprocessorBindGroup {
        pub camera: CameraPod, pub model_transform: ModelTransformPod, pub gaussian_transform: GaussianTransformPod,
        pub d_gaussians: *const c_void, pub gaussians_bytes: u64,
        pub d_indirect_args: *mut DrawIndirectArgs, pub d_radix_sort_indirect_args: *mut DispatchIndirectArgs,
        pub d_indirect_indices: *mut u32, pub indirect_indices_bytes: u64,
        pub d_gaussians_depth: *mut f32, pub gaussians_depth_bytes: u64,
        pub d_selection: *const u32, pub invert_selection: u32,
    }
    /// The `Renderer` bind group (reference src/renderer.rs:56-116): bindings 0-4.
    #[repr(C)] #[derive(Clone, Copy)]
    pub struct RendererBindGroup {
        pub camera: CameraPod, pub model_transform: ModelTransformPod, pub gaussian_transform: GaussianTransformPod,
        pub d_gaussians: *const c_void, pub gaussians_bytes: u64,
        pub d_indirect_indices: *const u32, pub indirect_indices_bytes: u64,
    }

    #[link(name = "splat_b200")]
    extern "C" {
        pub fn sb_preprocessor_create(ctx: *mut SbContext, sh: i32, cov: i32, n: u64, out: *mut *mut SbPreprocessor) -> i32;
        pub fn sb_preprocessor_destroy(p: *mut SbPreprocessor);
        pub fn sb_preprocessor_preprocess(p: *mut SbPreprocessor, stream: *mut c_void, bind_group: *const PreprocessorBindGroup, gaussian_count: u32) -> i32;
        pub fn sb_sorter_create(ctx: *mut SbContext, capacity: u32, out: *mut *mut SbRadixSorter) -> i32;
        pub fn sb_sorter_destroy(s: *mut SbRadixSorter);
        pub fn sb_sorter_sort(s: *mut SbRadixSorter, stream: *mut c_void, d_keys: *mut u32, d_payload: *mut u32, d_count: *const u32, max_count: u32, begin_bit: i32, end_bit: i32) -> i32;
        pub fn sb_renderer_create(ctx: *mut SbContext, sh: i32, cov: i32, target_format: i32, n: u64, out: *mut *mut SbRenderer) -> i32;
        pub fn sb_renderer_destroy(r: *mut SbRenderer);
        pub fn sb_renderer_render(r: *mut SbRenderer, stream: *mut c_void, bind_group: *const RendererBindGroup, target: *const Target, d_indirect_args: *const DrawIndirectArgs, depth: *const DepthAttachment, load: i32) -> i32;
        pub fn sb_last_error_string(ctx: *const SbContext) -> *const c_char;
        pub fn sb_pod_stride(sh_fmt: i32, cov_fmt: i32) -> u32;
        pub fn sb_pack_gaussians(src: *const Gaussian, n: u64, sh_fmt: i32, cov_fmt: i32, out: *mut c_void) -> i32;
        pub fn sb_camera_pod(pos: *const f32, yaw: f32, pitch: f32, z_near: f32, z_far: f32, fov: f32, w: u32, h: u32, out: *mut CameraPod) -> i32;
        pub fn sb_ctx_create(device: i32, out: *mut *mut SbContext) -> i32;
        pub fn sb_ctx_destroy(ctx: *mut SbContext);
        pub fn sb_viewer_create_from_gaussians(ctx: *mut SbContext, sh: i32, cov: i32, target_format: i32, src: *const Gaussian, n: u64, out: *mut *mut SbViewer) -> i32;
        pub fn sb_viewer_create(ctx: *mut SbContext, sh: i32, cov: i32, target_format: i32, pods: *const c_void, n: u64, out: *mut *mut SbViewer) -> i32;
        pub fn sb_viewer_destroy(v: *mut SbViewer);
        pub fn sb_viewer_update_camera_with_pod(v: *mut SbViewer, pod: *const CameraPod) -> i32;
        pub fn sb_viewer_update_model_transform_with_pod(v: *mut SbViewer, pod: *const ModelTransformPod) -> i32;
        pub fn sb_viewer_update_gaussian_transform_with_pod(v: *mut SbViewer, pod: *const GaussianTransformPod) -> i32;
        pub fn sb_viewer_enable_selection(v: *mut SbViewer, enabled: i32) -> i32;
        pub fn sb_viewer_set_selection(v: *mut SbViewer, stream: *mut c_void, words: *const u32, n_words: u64) -> i32;
        pub fn sb_viewer_set_invert_selection(v: *mut SbViewer, invert: i32) -> i32;
        pub fn sb_viewer_select_rect(v: *mut SbViewer, stream: *mut c_void, x0: f32, y0: f32, x1: f32, y1: f32) -> i32;
        pub fn sb_viewer_select_brush(v: *mut SbViewer, stream: *mut c_void, points_xy: *const f32, n_points: u32, radius: f32, accumulate: i32) -> i32;
        pub fn sb_viewer_apply_rgb_override(v: *mut SbViewer, stream: *mut c_void, rgb: *const f32, alpha: f32) -> i32;
        pub fn sb_viewer_restore_gaussians(v: *mut SbViewer, stream: *mut c_void) -> i32;
        pub fn sb_viewer_set_exact_cutoff(v: *mut SbViewer, enabled: i32) -> i32;
        pub fn sb_viewer_render(v: *mut SbViewer, stream: *mut c_void, target: *const Target) -> i32;
        pub fn sb_viewer_render_with_pass(v: *mut SbViewer, stream: *mut c_void, target: *const Target, depth: *const DepthAttachment, load: i32, run_stages: i32) -> i32;
        pub fn sb_viewer_preprocess(v: *mut SbViewer, stream: *mut c_void) -> i32;
        pub fn sb_viewer_sort(v: *mut SbViewer, stream: *mut c_void) -> i32;
        pub fn sb_viewer_draw(v: *mut SbViewer, stream: *mut c_void, target: *const Target) -> i32;
        pub fn sb_viewer_read_indirect_args(v: *mut SbViewer, stream: *mut c_void, draw: *mut DrawIndirectArgs, dispatch: *mut DispatchIndirectArgs) -> i32;
        pub fn sb_mm_create(ctx: *mut SbContext, sh: i32, cov: i32, target_format: i32, out: *mut *mut SbMultiModelViewer) -> i32;
        pub fn sb_mm_destroy(mm: *mut SbMultiModelViewer);
        pub fn sb_mm_insert_model(mm: *mut SbMultiModelViewer, key: u64, pods: *const c_void, n: u64, replaced: *mut i32) -> i32;
        pub fn sb_mm_remove_model(mm: *mut SbMultiModelViewer, key: u64, removed: *mut i32) -> i32;
        pub fn sb_mm_update_camera_with_pod(mm: *mut SbMultiModelViewer, pod: *const CameraPod) -> i32;
        pub fn sb_mm_update_model_transform_with_pod(mm: *mut SbMultiModelViewer, key: u64, pod: *const ModelTransformPod) -> i32;
        pub fn sb_mm_update_gaussian_transform_with_pod(mm: *mut SbMultiModelViewer, pod: *const GaussianTransformPod) -> i32;
        pub fn sb_mm_render(mm: *mut SbMultiModelViewer, stream: *mut c_void, target: *const Target, keys: *const u64, n_keys: u32) -> i32;
    }
}

pub use ffi::{CameraPod, Gaussian, GaussianTransformPod, ModelTransformPod, Target};

/// Mirrors `ViewerCreateError` / `MultiModelViewerAccessError` (reference src/error.rs:7-50).
#[derive(Debug, thiserror::Error)]
pub enum Error {
    #[error("model size exceeds the device limit: {0}")] ModelSizeExceedsDeviceLimit(String),
    #[error("model not found")] ModelNotFound,
    #[error("buffer size mismatch: {0}")] BadBufferSize(String),
    #[error("invalid argument: {0}")] InvalidArg(String),
    #[error("CUDA error: {0}")] Cuda(String),
}

fn check(status: i32, ctx: *const ffi::SbContext) -> Result<(), Error> {
    if status == 0 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(ffi::sb_last_error_string(ctx)) }.to_string_lossy().into_owned();
    Err(match status { 3 => Error::ModelSizeExceedsDeviceLimit(msg), 4 => Error::ModelNotFound, 5 => Error::BadBufferSize(msg), 1 => Error::InvalidArg(msg), _ => Error::Cuda(msg) })
}

/// The pod format marker trait, as `core::GaussianPod` (compile-time format → kernel instantiation).
pub trait GaussianPod { const SH: i32; const COV: i32; }
pub struct GaussianPodWithShSingleCov3dSingleConfigs; impl GaussianPod for GaussianPodWithShSingleCov3dSingleConfigs { const SH: i32 = 0; const COV: i32 = 0; }
pub struct GaussianPodWithShHalfCov3dHalfConfigs;     impl GaussianPod for GaussianPodWithShHalfCov3dHalfConfigs     { const SH: i32 = 1; const COV: i32 = 1; }
pub struct GaussianPodWithShNorm8Cov3dHalfConfigs;    impl GaussianPod for GaussianPodWithShNorm8Cov3dHalfConfigs    { const SH: i32 = 2; const COV: i32 = 1; }
pub type DefaultGaussianPod = GaussianPodWithShSingleCov3dSingleConfigs; // src/lib.rs:44

/// `CameraTrait` + `Camera` exactly as the reference (src/camera.rs:8-93).
pub trait CameraTrait { fn view(&self) -> glam::Mat4; fn projection(&self, aspect_ratio: f32) -> glam::Mat4; }

impl CameraPod {
    /// `CameraPod::new(camera, size)` — src/buffer/camera.rs:72-80.
    pub fn new(camera: &impl CameraTrait, size: glam::UVec2) -> Self {
        Self { view: camera.view().to_cols_array(), proj: camera.projection(size.x as f32 / size.y as f32).to_cols_array(), size: [size.x as f32, size.y as f32], _padding: [0; 2] }
    }
}

/// Replaces `&wgpu::Device`.
pub struct Context(*mut ffi::SbContext);
impl Context {
    pub fn new(device_ordinal: i32) -> Result<Self, Error> { let mut p = std::ptr::null_mut(); check(unsafe { ffi::sb_ctx_create(device_ordinal, &mut p) }, std::ptr::null())?; Ok(Self(p)) }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::sb_ctx_destroy(self.0) } } }

/// A CUDA stream handle (`cudaStream_t`); `Stream::DEFAULT` is the legacy default stream.
#[derive(Clone, Copy)] pub struct Stream(pub *mut c_void);
impl Stream { pub const DEFAULT: Stream = Stream(std::ptr::null_mut()); }

/// `wgpu_3dgs_viewer::Viewer<G>` (src/lib.rs:65-276).
pub struct Viewer<'c, G: GaussianPod = DefaultGaussianPod> { raw: *mut ffi::SbViewer, ctx: &'c Context, _g: PhantomData<G> }

impl<'c, G: GaussianPod> Viewer<'c, G> {
    /// `Viewer::new(device, texture_format, gaussians)`.
    pub fn new(ctx: &'c Context, texture_format: i32, gaussians: &[Gaussian]) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::sb_viewer_create_from_gaussians(ctx.0, G::SH, G::COV, texture_format, gaussians.as_ptr(), gaussians.len() as u64, &mut raw) }, ctx.0)?;
        Ok(Self { raw, ctx, _g: PhantomData })
    }
    pub fn update_camera(&mut self, camera: &impl CameraTrait, texture_size: glam::UVec2) { self.update_camera_with_pod(&CameraPod::new(camera, texture_size)) }
    pub fn update_camera_with_pod(&mut self, pod: &CameraPod) { unsafe { ffi::sb_viewer_update_camera_with_pod(self.raw, pod) }; }
    pub fn update_model_transform(&mut self, pos: glam::Vec3, rot: glam::Quat, scale: glam::Vec3) {
        self.update_model_transform_with_pod(&ModelTransformPod { pos: pos.to_array(), _pad0: 0.0, rot: rot.to_array(), scale: scale.to_array(), _pad1: 0.0 })
    }
    pub fn update_model_transform_with_pod(&mut self, pod: &ModelTransformPod) { unsafe { ffi::sb_viewer_update_model_transform_with_pod(self.raw, pod) }; }
    pub fn update_gaussian_transform_with_pod(&mut self, pod: &GaussianTransformPod) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_update_gaussian_transform_with_pod(self.raw, pod) }, self.ctx.0) }
    /// `Viewer::render(encoder, texture_view)`: enqueue preprocess → sort → draw on `stream`.
    pub fn render(&self, stream: Stream, target: &Target) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_render(self.raw, stream.0, target) }, self.ctx.0) }
    /// The three public stages (`viewer.preprocessor / radix_sorter / renderer`).
    pub fn preprocess(&self, stream: Stream) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_preprocess(self.raw, stream.0) }, self.ctx.0) }
    pub fn sort(&self, stream: Stream) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_sort(self.raw, stream.0) }, self.ctx.0) }
    pub fn draw(&self, stream: Stream, target: &Target) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_draw(self.raw, stream.0, target) }, self.ctx.0) }
    /// `Renderer::render_with_pass(pass, indirect_args)` (src/renderer.rs:187-195): draw inside the caller's pass —
    /// over the colour the target already holds, depth-tested against the pass's depth attachment when the viewer
    /// was created with `ViewerCreateOptions { depth_stencil: Some(..) }`.
    pub fn render_with_pass(&self, stream: Stream, target: &Target, depth: Option<&ffi::DepthAttachment>) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_render_with_pass(self.raw, stream.0, target, depth.map_or(std::ptr::null(), |d| d as *const _), 1, 0) }, self.ctx.0)
    }
    /// `selection::ViewportSelector` evaluation with a rectangle / brush mask (src/selection/viewport_selector.rs:250-301).
    pub fn select_rect(&mut self, stream: Stream, min: glam::Vec2, max: glam::Vec2) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_select_rect(self.raw, stream.0, min.x, min.y, max.x, max.y) }, self.ctx.0)
    }
    pub fn select_brush(&mut self, stream: Stream, stroke: &[glam::Vec2], radius: f32, accumulate: bool) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_select_brush(self.raw, stream.0, stroke.as_ptr().cast(), stroke.len() as u32, radius, accumulate as i32) }, self.ctx.0)
    }
}
impl<G: GaussianPod> Drop for Viewer<'_, G> { fn drop(&mut self) { unsafe { ffi::sb_viewer_destroy(self.raw) } } }

/// `MultiModelViewer<G, K>` (src/multi_model.rs:291-531); keys are hashed to u64 by the caller.
pub struct MultiModelViewer<'c, G: GaussianPod = DefaultGaussianPod> { raw: *mut ffi::SbMultiModelViewer, ctx: &'c Context, _g: PhantomData<G> }
impl<'c, G: GaussianPod> MultiModelViewer<'c, G> {
    pub fn new(ctx: &'c Context, texture_format: i32) -> Result<Self, Error> { let mut raw = std::ptr::null_mut(); check(unsafe { ffi::sb_mm_create(ctx.0, G::SH, G::COV, texture_format, &mut raw) }, ctx.0)?; Ok(Self { raw, ctx, _g: PhantomData }) }
    pub fn insert_model(&mut self, key: u64, gaussians: &[Gaussian]) -> Result<bool, Error> {
        let stride = unsafe { ffi::sb_pod_stride(G::SH, G::COV) } as usize;
        let mut pods = vec![0u8; stride * gaussians.len()];
        check(unsafe { ffi::sb_pack_gaussians(gaussians.as_ptr(), gaussians.len() as u64, G::SH, G::COV, pods.as_mut_ptr().cast()) }, self.ctx.0)?;
        let mut replaced = 0;
        check(unsafe { ffi::sb_mm_insert_model(self.raw, key, pods.as_ptr().cast(), gaussians.len() as u64, &mut replaced) }, self.ctx.0)?;
        Ok(replaced != 0)
    }
    pub fn remove_model(&mut self, key: u64) -> bool { let mut r = 0; unsafe { ffi::sb_mm_remove_model(self.raw, key, &mut r) }; r != 0 }
    pub fn update_camera_with_pod(&mut self, pod: &CameraPod) { unsafe { ffi::sb_mm_update_camera_with_pod(self.raw, pod) }; }
    pub fn update_model_transform_with_pod(&mut self, key: u64, pod: &ModelTransformPod) -> Result<(), Error> { check(unsafe { ffi::sb_mm_update_model_transform_with_pod(self.raw, key, pod) }, self.ctx.0) }
    pub fn update_gaussian_transform_with_pod(&mut self, pod: &GaussianTransformPod) -> Result<(), Error> { check(unsafe { ffi::sb_mm_update_gaussian_transform_with_pod(self.raw, pod) }, self.ctx.0) }
    /// `render(encoder, view, keys)` → `Err(ModelNotFound)` like the reference.
    pub fn render(&self, stream: Stream, target: &Target, keys: &[u64]) -> Result<(), Error> { check(unsafe { ffi::sb_mm_render(self.raw, stream.0, target, keys.as_ptr(), keys.len() as u32) }, self.ctx.0) }
}
impl<G: GaussianPod> Drop for MultiModelViewer<'_, G> { fn drop(&mut self) { unsafe { ffi::sb_mm_destroy(self.raw) } } }

/// `Preprocessor<G, ()>` (src/preprocessor.rs:370-450): `new_without_bind_group` + `preprocess(encoder, bind_group, n)`.
pub struct Preprocessor<'c, G: GaussianPod = DefaultGaussianPod> { raw: *mut ffi::SbPreprocessor, ctx: &'c Context, _g: PhantomData<G> }
impl<'c, G: GaussianPod> Preprocessor<'c, G> {
    pub fn new_without_bind_group(ctx: &'c Context, gaussian_capacity: u64) -> Result<Self, Error> { let mut raw = std::ptr::null_mut(); check(unsafe { ffi::sb_preprocessor_create(ctx.0, G::SH, G::COV, gaussian_capacity, &mut raw) }, ctx.0)?; Ok(Self { raw, ctx, _g: PhantomData }) }
    pub fn preprocess(&self, stream: Stream, bind_group: &ffi::PreprocessorBindGroup, gaussian_count: u32) -> Result<(), Error> { check(unsafe { ffi::sb_preprocessor_preprocess(self.raw, stream.0, bind_group, gaussian_count) }, self.ctx.0) }
}
impl<G: GaussianPod> Drop for Preprocessor<'_, G> { fn drop(&mut self) { unsafe { ffi::sb_preprocessor_destroy(self.raw) } } }

/// `RadixSorter<()>` (src/radix_sorter.rs:71-96): stable ascending sort of (depth key, index); the count is read on the
/// device from `IndirectArgsBuffer.instance_count`.
pub struct RadixSorter<'c> { raw: *mut ffi::SbRadixSorter, ctx: &'c Context }
impl<'c> RadixSorter<'c> {
    pub fn new_without_bind_groups(ctx: &'c Context, capacity: u32) -> Result<Self, Error> { let mut raw = std::ptr::null_mut(); check(unsafe { ffi::sb_sorter_create(ctx.0, capacity, &mut raw) }, ctx.0)?; Ok(Self { raw, ctx }) }
    pub fn sort(&self, stream: Stream, d_depth_keys: *mut f32, d_indices: *mut u32, d_indirect_args: *const ffi::DrawIndirectArgs, capacity: u32) -> Result<(), Error> {
        let d_count = unsafe { std::ptr::addr_of!((*d_indirect_args).instance_count) };
        check(unsafe { ffi::sb_sorter_sort(self.raw, stream.0, d_depth_keys.cast(), d_indices, d_count, capacity, 0, 32) }, self.ctx.0)
    }
}
impl Drop for RadixSorter<'_> { fn drop(&mut self) { unsafe { ffi::sb_sorter_destroy(self.raw) } } }

/// `Renderer<G, ()>` (src/renderer.rs:242-356): `render` clears to BLACK, `render_with_pass` composites over the target.
pub struct Renderer<'c, G: GaussianPod = DefaultGaussianPod> { raw: *mut ffi::SbRenderer, ctx: &'c Context, _g: PhantomData<G> }
impl<'c, G: GaussianPod> Renderer<'c, G> {
    pub fn new_without_bind_group(ctx: &'c Context, texture_format: i32, gaussian_capacity: u64) -> Result<Self, Error> { let mut raw = std::ptr::null_mut(); check(unsafe { ffi::sb_renderer_create(ctx.0, G::SH, G::COV, texture_format, gaussian_capacity, &mut raw) }, ctx.0)?; Ok(Self { raw, ctx, _g: PhantomData }) }
    pub fn render(&self, stream: Stream, target: &Target, bind_group: &ffi::RendererBindGroup, d_indirect_args: *const ffi::DrawIndirectArgs) -> Result<(), Error> { check(unsafe { ffi::sb_renderer_render(self.raw, stream.0, bind_group, target, d_indirect_args, std::ptr::null(), 0) }, self.ctx.0) }
    pub fn render_with_pass(&self, stream: Stream, target: &Target, depth: Option<&ffi::DepthAttachment>, bind_group: &ffi::RendererBindGroup, d_indirect_args: *const ffi::DrawIndirectArgs) -> Result<(), Error> { check(unsafe { ffi::sb_renderer_render(self.raw, stream.0, bind_group, target, d_indirect_args, depth.map_or(std::ptr::null(), |d| d as *const _), 1) }, self.ctx.0) }
}
impl<G: GaussianPod> Drop for Renderer<'_, G> { fn drop(&mut self) { unsafe { ffi::sb_renderer_destroy(self.raw) } } }
