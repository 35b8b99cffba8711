//! `splat-b200`: the per-frame splat pipeline of `wgpu-3dgs-viewer` on an NVIDIA B200, behind the
//! same shape of API (`Viewer::new / update_camera / render`, stage access, `MultiModelViewer`).
//!
//! Where the reference takes `&wgpu::Device`, `&wgpu::Queue`, `&mut wgpu::CommandEncoder` and a
//! `&wgpu::TextureView`, this crate takes a [`Context`], nothing, a CUDA stream and a [`Target`]
//! (caller-owned device memory).  `render` only enqueues work, like recording into an encoder.
//!
//! SOURCE ONLY — never compiled in the build container (no Rust toolchain there).  The `extern "C"` block is generated from
//! the header (all exports); the safe layer below wraps every one of them under the reference's names.
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
    #[repr(C)] pub struct SbStrips { _p: [u8; 0] }
    /// The three CUDA-IPC handles a rank publishes for the partitioned strips (records, tile boxes, inbox).
    #[repr(C)] #[derive(Clone, Copy)] pub struct StripsExport { pub handles: [[u8; 64]; 3] }

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

    /// editor `BasicColorModifiersPod` (tests/e2e/selection.rs:80-92): rgb override or HSV, then alpha / contrast / exposure / gamma.
    #[repr(C)] #[derive(Clone, Copy, Debug)]
    pub struct BasicColorModifiers { pub rgb_override: i32, pub rgb_or_hsv: [f32; 3], pub alpha: f32, pub contrast: f32, pub exposure: f32, pub gamma: f32 }
    impl Default for BasicColorModifiers { fn default() -> Self { Self { rgb_override: 0, rgb_or_hsv: [0.0, 1.0, 1.0], alpha: 1.0, contrast: 0.0, exposure: 0.0, gamma: 1.0 } } }

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

    // every export of include/splat_b200.h (generated: scripts/gen_rust_ffi.py; tests/test_abi.py keeps it in step with the header)
    include!("ffi_gen.rs");
}

/// `wgpu::TextureFormat`s the renderer can target (`SB_TARGET_*`): the reference takes any format (src/renderer.rs:120-155).
pub mod texture_format {
    pub const RGBA8_UNORM: i32 = 0; pub const BGRA8_UNORM: i32 = 1; pub const RGBA16_FLOAT: i32 = 2; pub const RGBA32_FLOAT: i32 = 3;
    pub const RGBA8_UNORM_SRGB: i32 = 4; pub const BGRA8_UNORM_SRGB: i32 = 5;
}
/// `GaussianDisplayMode` and `wgpu::CompareFunction` as the C ABI numbers them.
pub mod display_mode { pub const SPLAT: i32 = 0; pub const ELLIPSE: i32 = 1; pub const POINT: i32 = 2; }
pub mod compare { pub const NEVER: i32 = 1; pub const LESS: i32 = 2; pub const EQUAL: i32 = 3; pub const LESS_EQUAL: i32 = 4; pub const GREATER: i32 = 5; pub const NOT_EQUAL: i32 = 6; pub const GREATER_EQUAL: i32 = 7; pub const ALWAYS: i32 = 8; }

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
/// A device buffer of a [`Viewer`], as the reference's public buffer fields expose it (`viewer.gaussians_buffer`,
/// `.indirect_args_buffer`, `.radix_sort_indirect_args_buffer`, `.indirect_indices_buffer`, `.gaussians_depth_buffer`,
/// `.selection_buffer`: src/lib.rs:65-82): a raw device pointer + size, valid while the viewer lives.
#[derive(Clone, Copy, Debug)] pub struct DeviceBuffer<T> { pub ptr: *const T, pub len: u64 }

/// `ViewerCreateOptions` (src/lib.rs:279-293).
#[derive(Clone, Copy, Debug, Default)] pub struct ViewerCreateOptions { pub depth_stencil: Option<(i32 /* compare */, bool /* depth write */)> }

impl<'c, G: GaussianPod> Viewer<'c, G> {
    /// `Viewer::new_with_options` — the depth-stencil state is passed per pass here (`render_with_pass`), so the options only record it.
    pub fn new_with_options(ctx: &'c Context, texture_format: i32, gaussians: &[Gaussian], _options: ViewerCreateOptions) -> Result<Self, Error> { Self::new(ctx, texture_format, gaussians) }
    /// From pods already packed by `wgpu-3dgs-core` (`GaussiansBuffer::<G>` contents), host or device resident.
    pub fn from_pods(ctx: &'c Context, texture_format: i32, pods: &[u8]) -> Result<Self, Error> {
        let n = pods.len() as u64 / unsafe { ffi::sb_pod_stride(G::SH, G::COV) } as u64;
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::sb_viewer_create(ctx.0, G::SH, G::COV, texture_format, pods.as_ptr().cast(), n, &mut raw) }, ctx.0)?;
        Ok(Self { raw, ctx, _g: PhantomData })
    }
    /// `TryFrom<wgpu::Buffer>`: adopt a caller-owned device buffer after the size check (src/buffer/camera.rs:42-58 pattern).
    pub unsafe fn from_device(ctx: &'c Context, texture_format: i32, d_pods: *const c_void, bytes: u64, n: u64) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(ffi::sb_viewer_create_from_device(ctx.0, G::SH, G::COV, texture_format, d_pods, bytes, n, &mut raw), ctx.0)?;
        Ok(Self { raw, ctx, _g: PhantomData })
    }
    pub fn update_camera_pose(&mut self, pos: glam::Vec3, yaw: f32, pitch: f32, z: std::ops::Range<f32>, vertical_fov: f32, size: glam::UVec2) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_update_camera(self.raw, pos.to_array().as_ptr(), yaw, pitch, z.start, z.end, vertical_fov, size.x, size.y) }, self.ctx.0)
    }
    pub fn update_model_transform_raw(&mut self, pos: glam::Vec3, rot: glam::Quat, scale: glam::Vec3) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_update_model_transform(self.raw, pos.to_array().as_ptr(), rot.to_array().as_ptr(), scale.to_array().as_ptr()) }, self.ctx.0)
    }
    /// `update_gaussian_transform(queue, size, display_mode, sh_deg, no_sh0, max_std_dev)` (src/lib.rs:237-254).
    pub fn update_gaussian_transform(&mut self, size: f32, display_mode: i32, sh_deg: i32, no_sh0: bool, max_std_dev: f32) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_update_gaussian_transform(self.raw, size, display_mode, sh_deg, no_sh0 as i32, max_std_dev) }, self.ctx.0)
    }
    // ---- the public buffer fields of the reference Viewer
    pub fn gaussians_buffer(&self) -> Result<DeviceBuffer<c_void>, Error> { let (mut p, mut b) = (std::ptr::null(), 0u64); check(unsafe { ffi::sb_viewer_gaussians_ptr(self.raw, &mut p, &mut b) }, self.ctx.0)?; Ok(DeviceBuffer { ptr: p, len: b }) }
    pub fn indirect_args_buffer(&self) -> Result<DeviceBuffer<ffi::DrawIndirectArgs>, Error> { let mut p = std::ptr::null(); check(unsafe { ffi::sb_viewer_indirect_args_ptr(self.raw, &mut p) }, self.ctx.0)?; Ok(DeviceBuffer { ptr: p, len: 16 }) }
    pub fn radix_sort_indirect_args_buffer(&self) -> Result<DeviceBuffer<ffi::DispatchIndirectArgs>, Error> { let mut p = std::ptr::null(); check(unsafe { ffi::sb_viewer_radix_sort_indirect_args_ptr(self.raw, &mut p) }, self.ctx.0)?; Ok(DeviceBuffer { ptr: p, len: 12 }) }
    pub fn indirect_indices_buffer(&self) -> Result<DeviceBuffer<u32>, Error> { let (mut p, mut c) = (std::ptr::null(), 0u64); check(unsafe { ffi::sb_viewer_indirect_indices_ptr(self.raw, &mut p, &mut c) }, self.ctx.0)?; Ok(DeviceBuffer { ptr: p, len: c }) }
    pub fn gaussians_depth_buffer(&self) -> Result<DeviceBuffer<f32>, Error> { let (mut p, mut b) = (std::ptr::null(), 0u64); check(unsafe { ffi::sb_viewer_gaussians_depth_ptr(self.raw, &mut p, &mut b) }, self.ctx.0)?; Ok(DeviceBuffer { ptr: p, len: b }) }
    /// `viewer.selection_buffer` (editor `SelectionBuffer`: bit i%32 of word i/32) — enable the feature first.
    pub fn selection_buffer(&self) -> Result<DeviceBuffer<u32>, Error> { let (mut p, mut n) = (std::ptr::null_mut(), 0u64); check(unsafe { ffi::sb_viewer_selection_ptr(self.raw, &mut p, &mut n) }, self.ctx.0)?; Ok(DeviceBuffer { ptr: p as *const u32, len: n }) }
    pub fn enable_selection(&mut self, enabled: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_enable_selection(self.raw, enabled as i32) }, self.ctx.0) }
    pub fn set_selection(&mut self, stream: Stream, words: &[u32]) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_selection(self.raw, stream.0, words.as_ptr(), words.len() as u64) }, self.ctx.0) }
    pub fn read_selection(&self, stream: Stream, out: &mut [u32]) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_read_selection(self.raw, stream.0, out.as_mut_ptr(), out.len() as u64) }, self.ctx.0) }
    /// `invert_selection_buffer.update(queue, invert)` (src/selection/buffer.rs:167-171); default = inverted.
    pub fn set_invert_selection(&mut self, invert: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_invert_selection(self.raw, invert as i32) }, self.ctx.0) }
    /// editor `NonDestructiveModifier<BasicSelectionModifier>` with an rgb override (tests/e2e/selection.rs:54-116).
    pub fn apply_rgb_override(&mut self, stream: Stream, rgb: glam::Vec3, alpha: f32) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_apply_rgb_override(self.raw, stream.0, rgb.to_array().as_ptr(), alpha) }, self.ctx.0) }
    /// `basic_color_modifiers_buffer.update_with_pod(queue, &BasicColorModifiersPod { .. })` + `modifier.apply(..)` (examples/selection.rs:218-229).
    pub fn apply_basic_color_modifiers(&mut self, stream: Stream, m: &ffi::BasicColorModifiers) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_apply_basic_color_modifiers(self.raw, stream.0, m) }, self.ctx.0) }
    pub fn restore_gaussians(&mut self, stream: Stream) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_restore_gaussians(self.raw, stream.0) }, self.ctx.0) }
    // ---- frames beyond `render`
    /// One frame straight to (pinned) host memory: camera pod in, pixels out.
    pub fn render_to_host(&mut self, stream: Stream, camera: &CameraPod, pixels: &mut [u8]) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_render_to_host(self.raw, stream.0, camera, pixels.as_mut_ptr().cast(), pixels.len() as u64) }, self.ctx.0) }
    /// A batch of camera views of the same scene (two in flight); `targets` are device frames.
    pub fn render_batch(&mut self, stream: Stream, cameras: &[CameraPod], targets: &[Target]) -> Result<(), Error> {
        assert_eq!(cameras.len(), targets.len());
        check(unsafe { ffi::sb_viewer_render_batch(self.raw, stream.0, cameras.as_ptr(), targets.as_ptr(), std::ptr::null(), cameras.len() as u32) }, self.ctx.0)
    }
    pub fn render_batch_to_host(&mut self, stream: Stream, cameras: &[CameraPod], host_frames: &[*mut c_void]) -> Result<(), Error> {
        assert_eq!(cameras.len(), host_frames.len());
        check(unsafe { ffi::sb_viewer_render_batch(self.raw, stream.0, cameras.as_ptr(), std::ptr::null(), host_frames.as_ptr(), cameras.len() as u32) }, self.ctx.0)
    }
    /// `render_with_pass` after running the stages too (one call per frame inside a caller's pass).
    pub fn render_in_pass(&self, stream: Stream, target: &Target, depth: Option<&ffi::DepthAttachment>, load: bool) -> Result<(), Error> {
        check(unsafe { ffi::sb_viewer_render_with_pass(self.raw, stream.0, target, depth.map_or(std::ptr::null(), |d| d as *const _), load as i32, 1) }, self.ctx.0)
    }
    // ---- artefact read-back (synchronising; the reference's tests download the buffers the same way)
    pub fn read_indirect_args(&self, stream: Stream) -> Result<(ffi::DrawIndirectArgs, ffi::DispatchIndirectArgs), Error> {
        let (mut d, mut s) = (ffi::DrawIndirectArgs::default(), ffi::DispatchIndirectArgs::default());
        check(unsafe { ffi::sb_viewer_read_indirect_args(self.raw, stream.0, &mut d, &mut s) }, self.ctx.0)?; Ok((d, s))
    }
    pub fn read_indices(&self, stream: Stream, out: &mut [u32]) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_read_indices(self.raw, stream.0, out.as_mut_ptr(), out.len() as u64) }, self.ctx.0) }
    pub fn read_depth_keys(&self, stream: Stream, out: &mut [f32]) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_read_depth_keys(self.raw, stream.0, out.as_mut_ptr(), out.len() as u64) }, self.ctx.0) }
    /// (visible, tile duplicates, overflowed) of the last frame.
    pub fn read_frame_stats(&self, stream: Stream) -> Result<(u64, u64, bool), Error> { let (mut v, mut d, mut o) = (0u64, 0u64, 0u32); check(unsafe { ffi::sb_viewer_read_frame_stats(self.raw, stream.0, &mut v, &mut d, &mut o) }, self.ctx.0)?; Ok((v, d, o != 0)) }
    pub fn read_tile_row_work(&self, stream: Stream, out: &mut [u64]) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_read_tile_row_work(self.raw, stream.0, out.as_mut_ptr(), out.len() as u32) }, self.ctx.0) }
    // ---- knobs without a reference counterpart
    pub fn raster_path_is_gather4(&self) -> Result<bool, Error> { let mut g = 0; check(unsafe { ffi::sb_viewer_raster_path(self.raw, &mut g) }, self.ctx.0)?; Ok(g != 0) }
    pub fn set_strict_exp(&mut self, strict: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_strict_exp(self.raw, strict as i32) }, self.ctx.0) }
    pub fn set_exact_cutoff(&mut self, enabled: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_exact_cutoff(self.raw, enabled as i32) }, self.ctx.0) }
    /// Strip renders keep only the splats of their own strip (one frame sharded over GPUs).
    pub fn set_strip_cull(&mut self, enabled: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_strip_cull(self.raw, enabled as i32) }, self.ctx.0) }
    pub fn reserve_duplicates(&mut self, capacity: u64) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_reserve_duplicates(self.raw, capacity) }, self.ctx.0) }
    pub fn set_stage_timing(&mut self, enabled: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_stage_timing(self.raw, enabled as i32) }, self.ctx.0) }
    /// ms of preprocess, depth sort, tile count+emit, tile sort, gather, raster of the last frame.
    pub fn read_stage_times(&self, stream: Stream) -> Result<[f32; 6], Error> { let mut ms = [0f32; 6]; check(unsafe { ffi::sb_viewer_read_stage_times(self.raw, stream.0, ms.as_mut_ptr()) }, self.ctx.0)?; Ok(ms) }
    pub fn set_raster_counting(&mut self, enabled: bool) -> Result<(), Error> { check(unsafe { ffi::sb_viewer_set_raster_counting(self.raw, enabled as i32) }, self.ctx.0) }
    pub fn read_raster_counters(&self, stream: Stream) -> Result<(u64, u64), Error> { let (mut a, mut e) = (0u64, 0u64); check(unsafe { ffi::sb_viewer_read_raster_counters(self.raw, stream.0, &mut a, &mut e) }, self.ctx.0)?; Ok((a, e)) }
    pub fn read_raster_warp_counters(&self, stream: Stream) -> Result<(u64, u64), Error> { let (mut a, mut e) = (0u64, 0u64); check(unsafe { ffi::sb_viewer_read_raster_warp_counters(self.raw, stream.0, &mut a, &mut e) }, self.ctx.0)?; Ok((a, e)) }
}
impl<G: GaussianPod> Drop for Viewer<'_, G> { fn drop(&mut self) { unsafe { ffi::sb_viewer_destroy(self.raw) } } }

// ---- free functions of the boundary
pub fn version() -> String { unsafe { CStr::from_ptr(ffi::sb_version()) }.to_string_lossy().into_owned() }
pub fn status_string(status: i32) -> String { unsafe { CStr::from_ptr(ffi::sb_status_string(status)) }.to_string_lossy().into_owned() }
/// `size_of::<G>()` of the pod format (tests/e2e/multi_model.rs:43-53 pins buffer size = count * size_of::<G>()).
pub fn pod_stride<G: GaussianPod>() -> u32 { unsafe { ffi::sb_pod_stride(G::SH, G::COV) } }
/// `GaussiansDepthBuffer` size for n Gaussians (src/buffer/depth.rs:5-21) and the padded key count (src/radix_sorter.rs:922-939).
pub fn keys_buffer_size_bytes(n: u32) -> u64 { unsafe { ffi::sb_keys_buffer_size_bytes(n) } }
pub fn padded_key_count(n: u32) -> u32 { unsafe { ffi::sb_padded_key_count(n) } }
/// `Gaussians::read_from_file(path, GaussiansSource::Ply)` (examples/simple.rs:157-160).
pub fn read_ply(path: &std::path::Path) -> Result<Vec<Gaussian>, Error> {
    let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).map_err(|e| Error::InvalidArg(e.to_string()))?;
    let (mut p, mut n) = (std::ptr::null_mut::<Gaussian>(), 0u64);
    check(unsafe { ffi::sb_read_ply(c.as_ptr(), &mut p, &mut n) }, std::ptr::null())?;
    let v = unsafe { std::slice::from_raw_parts(p, n as usize) }.to_vec();
    unsafe { ffi::sb_free(p.cast()) };
    Ok(v)
}
/// `Gaussians::read_from_file(path, GaussiansSource::Spz)`: Niantic .spz v2 / v3.
pub fn read_spz(path: &std::path::Path) -> Result<Vec<Gaussian>, Error> {
    let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).map_err(|e| Error::InvalidArg(e.to_string()))?;
    let (mut p, mut n) = (std::ptr::null_mut::<Gaussian>(), 0u64);
    check(unsafe { ffi::sb_read_spz(c.as_ptr(), &mut p, &mut n) }, std::ptr::null())?;
    let v = unsafe { std::slice::from_raw_parts(p, n as usize) }.to_vec();
    unsafe { ffi::sb_free(p.cast()) };
    Ok(v)
}
/// `Gaussians::read_from_file`: PLY first, then SPZ, as the reference's examples try them (examples/simple.rs:157-160).
pub fn read_from_file(path: &std::path::Path) -> Result<Vec<Gaussian>, Error> { read_ply(path).or_else(|_| read_spz(path)) }
impl ModelTransformPod {
    pub fn new(pos: glam::Vec3, rot: glam::Quat, scale: glam::Vec3) -> Self { let mut o = std::mem::MaybeUninit::<Self>::uninit(); unsafe { ffi::sb_model_transform_pod(pos.to_array().as_ptr(), rot.to_array().as_ptr(), scale.to_array().as_ptr(), o.as_mut_ptr()); o.assume_init() } }
}
impl GaussianTransformPod {
    pub fn new(size: f32, display_mode: i32, sh_deg: i32, no_sh0: bool, max_std_dev: f32) -> Self { let mut o = std::mem::MaybeUninit::<Self>::uninit(); unsafe { ffi::sb_gaussian_transform_pod(size, display_mode, sh_deg, no_sh0 as i32, max_std_dev, o.as_mut_ptr()); o.assume_init() } }
}
impl CameraPod {
    /// The reference `Camera` (pos, yaw, pitch, z range, vertical fov) straight to a pod (src/camera.rs:71-93 + src/buffer/camera.rs:72-80).
    pub fn from_pose(pos: glam::Vec3, yaw: f32, pitch: f32, z: std::ops::Range<f32>, vertical_fov: f32, size: glam::UVec2) -> Self { let mut o = std::mem::MaybeUninit::<Self>::uninit(); unsafe { ffi::sb_camera_pod(pos.to_array().as_ptr(), yaw, pitch, z.start, z.end, vertical_fov, size.x, size.y, o.as_mut_ptr()); o.assume_init() } }
}
impl Context {
    /// `device.limits().max_storage_buffer_binding_size` of the reference's size check (src/preprocessor.rs:239-246).
    pub fn set_model_size_limit(&mut self, bytes: u64) -> Result<(), Error> { check(unsafe { ffi::sb_ctx_set_model_size_limit(self.0, bytes) }, self.0) }
    /// Measured FP32-issue (lane-ops/s, FMA = 1) and shared-memory (bytes/s) peaks of this device.
    pub fn probe_peaks(&self, stream: Stream) -> Result<(f64, f64), Error> { let (mut a, mut b) = (0f64, 0f64); check(unsafe { ffi::sb_probe_peaks(self.0, stream.0, &mut a, &mut b) }, self.0)?; Ok((a, b)) }
}

/// One frame as screen strips with the Preprocessor's work partitioned over the ranks (`sb_strips_*`): `scatter`, a barrier across
/// the ranks, `render`, a second barrier before the next frame.
pub struct Strips<'v, 'c, G: GaussianPod> { raw: *mut ffi::SbStrips, viewer: &'v Viewer<'c, G>, pub exported: ffi::StripsExport }
impl<'v, 'c, G: GaussianPod> Strips<'v, 'c, G> {
    /// `bounds[r] = (row0, rows)` of every rank's strip (whole 16-pixel tile rows).
    pub fn new(viewer: &'v Viewer<'c, G>, rank: u32, bounds: &[(u32, u32)]) -> Result<Self, Error> {
        let (r0, rn): (Vec<u32>, Vec<u32>) = bounds.iter().cloned().unzip();
        let (mut raw, mut ex) = (std::ptr::null_mut(), ffi::StripsExport { handles: [[0u8; 64]; 3] });
        check(unsafe { ffi::sb_strips_create(viewer.raw, bounds.len() as u32, rank, r0.as_ptr(), rn.as_ptr(), &mut raw, &mut ex) }, viewer.ctx.0)?;
        Ok(Self { raw, viewer, exported: ex })
    }
    pub fn connect(&mut self, all_ranks: &[ffi::StripsExport]) -> Result<(), Error> { check(unsafe { ffi::sb_strips_connect(self.raw, all_ranks.as_ptr()) }, self.viewer.ctx.0) }
    pub fn scatter(&self, stream: Stream) -> Result<(), Error> { check(unsafe { ffi::sb_strips_scatter(self.raw, stream.0) }, self.viewer.ctx.0) }
    pub fn render(&self, stream: Stream, target: &Target) -> Result<(), Error> { check(unsafe { ffi::sb_strips_render(self.raw, stream.0, target) }, self.viewer.ctx.0) }
}
impl<G: GaussianPod> Drop for Strips<'_, '_, G> { fn drop(&mut self) { unsafe { ffi::sb_strips_destroy(self.raw) } } }

/// A frame other GPUs of the node render their strips into (CUDA IPC over NVLink): the owner creates and publishes `handle`.
pub struct SharedFrame<'c> { pub ptr: *mut c_void, pub handle: [u8; 64], owner: bool, ctx: &'c Context }
impl<'c> SharedFrame<'c> {
    pub fn create(ctx: &'c Context, bytes: u64) -> Result<Self, Error> { let (mut p, mut h) = (std::ptr::null_mut(), [0u8; 64]); check(unsafe { ffi::sb_shared_frame_create(ctx.0, bytes, &mut p, h.as_mut_ptr()) }, ctx.0)?; Ok(Self { ptr: p, handle: h, owner: true, ctx }) }
    pub fn open(ctx: &'c Context, handle: [u8; 64]) -> Result<Self, Error> { let mut p = std::ptr::null_mut(); check(unsafe { ffi::sb_shared_frame_open(ctx.0, handle.as_ptr(), &mut p) }, ctx.0)?; Ok(Self { ptr: p, handle, owner: false, ctx }) }
}
impl Drop for SharedFrame<'_> { fn drop(&mut self) { unsafe { if self.owner { ffi::sb_shared_frame_destroy(self.ctx.0, self.ptr) } else { ffi::sb_shared_frame_close(self.ctx.0, self.ptr) } }; } }

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
impl<'c, G: GaussianPod> MultiModelViewer<'c, G> {
    /// `new_with_options(depth_stencil)` (src/multi_model.rs:319-337): the state is supplied per pass (`render_with_pass`).
    pub fn new_with_options(ctx: &'c Context, texture_format: i32, _options: ViewerCreateOptions) -> Result<Self, Error> { Self::new(ctx, texture_format) }
    /// `insert_model_with(device, key, gaussians_buffer)` (src/multi_model.rs:366): adopt device-resident pods.
    pub unsafe fn insert_model_with(&mut self, key: u64, d_pods: *const c_void, bytes: u64, n: u64) -> Result<bool, Error> { let mut r = 0; check(ffi::sb_mm_insert_model_from_device(self.raw, key, d_pods, bytes, n, &mut r), self.ctx.0)?; Ok(r != 0) }
    pub fn insert_model_from_gaussians(&mut self, key: u64, gaussians: &[Gaussian]) -> Result<bool, Error> { let mut r = 0; check(unsafe { ffi::sb_mm_insert_model_from_gaussians(self.raw, key, gaussians.as_ptr(), gaussians.len() as u64, &mut r) }, self.ctx.0)?; Ok(r != 0) }
    /// Non-pod updates (src/multi_model.rs:398-466).
    pub fn update_camera(&mut self, pos: glam::Vec3, yaw: f32, pitch: f32, z: std::ops::Range<f32>, vertical_fov: f32, size: glam::UVec2) -> Result<(), Error> { check(unsafe { ffi::sb_mm_update_camera(self.raw, pos.to_array().as_ptr(), yaw, pitch, z.start, z.end, vertical_fov, size.x, size.y) }, self.ctx.0) }
    pub fn update_model_transform(&mut self, key: u64, pos: glam::Vec3, rot: glam::Quat, scale: glam::Vec3) -> Result<(), Error> { check(unsafe { ffi::sb_mm_update_model_transform(self.raw, key, pos.to_array().as_ptr(), rot.to_array().as_ptr(), scale.to_array().as_ptr()) }, self.ctx.0) }
    pub fn update_gaussian_transform(&mut self, size: f32, display_mode: i32, sh_deg: i32, no_sh0: bool, max_std_dev: f32) -> Result<(), Error> { check(unsafe { ffi::sb_mm_update_gaussian_transform(self.raw, size, display_mode, sh_deg, no_sh0 as i32, max_std_dev) }, self.ctx.0) }
    // per-model selection (each model has its own SelectionBuffer in the reference's multi-model example)
    pub fn enable_selection(&mut self, key: u64, enabled: bool, invert: bool) -> Result<(), Error> { check(unsafe { ffi::sb_mm_enable_selection(self.raw, key, enabled as i32, invert as i32) }, self.ctx.0) }
    pub fn set_selection(&mut self, key: u64, stream: Stream, words: &[u32], invert: bool) -> Result<(), Error> { check(unsafe { ffi::sb_mm_set_selection(self.raw, key, stream.0, words.as_ptr(), words.len() as u64, invert as i32) }, self.ctx.0) }
    pub fn select_rect(&mut self, key: u64, stream: Stream, min: glam::Vec2, max: glam::Vec2) -> Result<(), Error> { check(unsafe { ffi::sb_mm_select_rect(self.raw, key, stream.0, min.x, min.y, max.x, max.y) }, self.ctx.0) }
    pub fn select_brush(&mut self, key: u64, stream: Stream, stroke: &[glam::Vec2], radius: f32, accumulate: bool) -> Result<(), Error> { check(unsafe { ffi::sb_mm_select_brush(self.raw, key, stream.0, stroke.as_ptr().cast(), stroke.len() as u32, radius, accumulate as i32) }, self.ctx.0) }
    pub fn read_selection(&self, key: u64, stream: Stream, out: &mut [u32]) -> Result<(), Error> { check(unsafe { ffi::sb_mm_read_selection(self.raw, key, stream.0, out.as_mut_ptr(), out.len() as u64) }, self.ctx.0) }
    /// `render_with_pass`-style multi-model frame: composite over the target, depth-tested against the pass's attachment.
    pub fn render_with_pass(&self, stream: Stream, target: &Target, depth: Option<&ffi::DepthAttachment>, load: bool, keys: &[u64]) -> Result<(), Error> { check(unsafe { ffi::sb_mm_render_with_pass(self.raw, stream.0, target, depth.map_or(std::ptr::null(), |d| d as *const _), load as i32, keys.as_ptr(), keys.len() as u32) }, self.ctx.0) }
    /// Sorted indices of one model after a frame, with its `IndirectArgsBuffer` contents ({6, V, 0, 0}).
    pub fn read_model_indices(&self, key: u64, stream: Stream, out: &mut [u32]) -> Result<ffi::DrawIndirectArgs, Error> { let mut d = ffi::DrawIndirectArgs::default(); check(unsafe { ffi::sb_mm_read_model_indices(self.raw, key, stream.0, out.as_mut_ptr(), out.len() as u64, &mut d) }, self.ctx.0)?; Ok(d) }
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
impl<G: GaussianPod> Renderer<'_, G> { pub fn set_strict_exp(&mut self, strict: bool) -> Result<(), Error> { check(unsafe { ffi::sb_renderer_set_strict_exp(self.raw, strict as i32) }, self.ctx.0) } }
impl<G: GaussianPod> Drop for Renderer<'_, G> { fn drop(&mut self) { unsafe { ffi::sb_renderer_destroy(self.raw) } } }
