//! A CUDA backend for Pathfinder 3's D3D11-level pipeline. It plugs in at the renderer protocol
//! (`begin_scene` / `render_command` / `end_scene`, renderer/src/gpu/renderer.rs:350-460) rather than
//! at `trait Device`, which is shaped around GLSL programs and raster draws.
//!
//! One change to the reference is needed to build this outside `pathfinder_renderer`: its `gpu_data`
//! module (RenderCommand and the batch records) is private (`renderer/src/lib.rs:28`); make it
//! `pub mod gpu_data;`, or move this file in next to `gpu/d3d11/renderer.rs` and `crate::` the paths.

mod ffi;

use pathfinder_color::ColorF;
use pathfinder_geometry::rect::RectF;
use pathfinder_geometry::vector::Vector2I;
use pathfinder_renderer::concurrent::executor::Executor;
use pathfinder_renderer::gpu::options::{RendererLevel, RendererMode};
use pathfinder_renderer::gpu_data::{PathSource, RenderCommand, SegmentsD3D11, TileBatchDataD3D11};
use pathfinder_renderer::options::{BuildOptions, RenderCommandListener};
use pathfinder_renderer::scene::{Scene, SceneSink};
use pathfinder_resources::ResourceLoader;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;
use std::sync::{Arc, Mutex};

pub struct CudaRenderer {
    raw: *mut c_void,
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::PFCudaGetLastError()).to_string_lossy().into_owned() }
}

fn check(status: i32) {
    // The reference panics on misuse (e.g. Renderer::require_d3d11, gpu/renderer.rs:1349-1360).
    if status != ffi::PF_CUDA_OK {
        panic!("pathfinder_cuda: status {}: {}", status, last_error());
    }
}

fn segments(s: &SegmentsD3D11) -> ffi::PFSegmentsD3D11 {
    ffi::PFSegmentsD3D11 {
        points: s.points.as_ptr(),
        point_count: s.points.len(),
        indices: s.indices.as_ptr(),
        index_count: s.indices.len(),
    }
}

fn batch(b: &TileBatchDataD3D11) -> ffi::PFTileBatchDataD3D11 {
    let t = &b.prepare_info.transform;
    let clipped = b.clipped_path_info.as_ref();
    ffi::PFTileBatchDataD3D11 {
        batch_id: b.batch_id.0,
        path_count: b.path_count,
        tile_count: b.tile_count,
        segment_count: b.segment_count,
        prepare_info: ffi::PFPrepareTilesInfoD3D11 {
            backdrops: b.prepare_info.backdrops.as_ptr(),
            backdrop_count: b.prepare_info.backdrops.len(),
            propagate_metadata: b.prepare_info.propagate_metadata.as_ptr(),
            dice_metadata: b.prepare_info.dice_metadata.as_ptr(),
            tile_path_info: b.prepare_info.tile_path_info.as_ptr(),
            transform: [t.matrix.m11(), t.matrix.m12(), t.matrix.m21(), t.matrix.m22(), t.vector.x(), t.vector.y()],
        },
        path_source: match b.path_source { PathSource::Draw => 0, PathSource::Clip => 1 },
        has_clipped_path_info: clipped.is_some() as u32,
        clipped_path_info: ffi::PFClippedPathInfo {
            clip_batch_id: clipped.map_or(0, |c| c.clip_batch_id.0),
            clipped_path_count: clipped.map_or(0, |c| c.clipped_path_count),
            max_clipped_tile_count: clipped.map_or(0, |c| c.max_clipped_tile_count),
        },
        content_key: 0,
    }
}

fn simple(kind: u32) -> ffi::PFRenderCommand {
    ffi::PFRenderCommand { kind, u: ffi::PFRenderCommandPayload { push_render_target: 0 } }
}

impl CudaRenderer {
    pub fn new(ordinal: i32, resources: &dyn ResourceLoader, mode: RendererMode, dest_size: Vector2I,
               background: Option<ColorF>) -> CudaRenderer {
        assert_eq!(mode.level, RendererLevel::D3D11);
        // The same resource Renderer::new loads (gpu/renderer.rs:207-222).
        let area = image::load_from_memory(&resources.slurp("textures/area-lut.png").unwrap()).unwrap().to_rgba();
        let options = ffi::PFCudaRendererOptions {
            dest_size: [dest_size.x(), dest_size.y()],
            background_color: background.map_or([0.0; 4], |c| [c.r(), c.g(), c.b(), c.a()]),
            flags: background.is_some() as u8,
        };
        let raw = unsafe {
            ffi::PFCudaRendererCreate(ffi::PFCudaDeviceCreate(ordinal), area.as_ptr(), ptr::null(),
                                      &ffi::PFRendererMode { level: 2 }, &options)
        };
        assert!(!raw.is_null(), "{}", last_error());
        CudaRenderer { raw }
    }

    pub fn set_view_box(&mut self, view_box: RectF) {
        check(unsafe { ffi::PFCudaRendererSetViewBox(self.raw, &view_box) })
    }

    pub fn begin_scene(&mut self) {
        check(unsafe { ffi::PFCudaRendererBeginScene(self.raw) })
    }

    pub fn end_scene(&mut self) {
        check(unsafe { ffi::PFCudaRendererEndScene(self.raw) })
    }

    /// Payloads are borrowed for the call; the renderer copies what it keeps.
    pub fn render_command(&mut self, command: &RenderCommand) {
        let c = match *command {
            RenderCommand::Start { path_count, needs_readable_framebuffer, .. } => ffi::PFRenderCommand {
                kind: ffi::START,
                u: ffi::PFRenderCommandPayload {
                    start: ffi::PFStart { path_count: path_count as u64,
                                          needs_readable_framebuffer: needs_readable_framebuffer as u32 },
                },
            },
            RenderCommand::UploadTextureMetadata(ref entries) => ffi::PFRenderCommand {
                kind: ffi::UPLOAD_TEXTURE_METADATA,
                u: ffi::PFRenderCommandPayload {
                    upload_texture_metadata: ffi::PFUploadTextureMetadata {
                        entries: entries.as_ptr(), entry_count: entries.len(), content_key: 0,
                    },
                },
            },
            RenderCommand::UploadSceneD3D11 { ref draw_segments, ref clip_segments } => ffi::PFRenderCommand {
                kind: ffi::UPLOAD_SCENE_D3D11,
                u: ffi::PFRenderCommandPayload {
                    upload_scene_d3d11: ffi::PFUploadSceneD3D11 {
                        draw_segments: segments(draw_segments),
                        clip_segments: segments(clip_segments),
                        payload_persists: 0,
                    },
                },
            },
            RenderCommand::PrepareClipTilesD3D11(ref b) => ffi::PFRenderCommand {
                kind: ffi::PREPARE_CLIP_TILES_D3D11,
                u: ffi::PFRenderCommandPayload { prepare_clip_tiles_d3d11: batch(b) },
            },
            RenderCommand::DrawTilesD3D11(ref draw) => ffi::PFRenderCommand {
                kind: ffi::DRAW_TILES_D3D11,
                u: ffi::PFRenderCommandPayload {
                    draw_tiles_d3d11: ffi::PFDrawTilesD3D11 {
                        tile_batch_data: batch(&draw.tile_batch_data),
                        has_color_texture: draw.color_texture.is_some() as u32,
                    },
                },
            },
            RenderCommand::PushRenderTarget(id) => ffi::PFRenderCommand {
                kind: ffi::PUSH_RENDER_TARGET,
                u: ffi::PFRenderCommandPayload { push_render_target: id.render_target },
            },
            RenderCommand::PopRenderTarget => simple(ffi::POP_RENDER_TARGET),
            RenderCommand::Finish { cpu_build_time } => ffi::PFRenderCommand {
                kind: ffi::FINISH,
                u: ffi::PFRenderCommandPayload { finish_cpu_build_time_ns: cpu_build_time.as_nanos() as u64 },
            },
            // D3D9-level commands are refused with PF_CUDA_ERROR_WRONG_LEVEL, textures / render targets
            // with PF_CUDA_ERROR_UNSUPPORTED: `check` turns both into a panic, like the reference.
            RenderCommand::AddFillsD3D9(_) => simple(ffi::ADD_FILLS_D3D9),
            RenderCommand::FlushFillsD3D9 => simple(ffi::FLUSH_FILLS_D3D9),
            RenderCommand::DrawTilesD3D9(_) => simple(ffi::DRAW_TILES_D3D9),
            RenderCommand::AllocateTexturePage { .. } => simple(ffi::ALLOCATE_TEXTURE_PAGE),
            RenderCommand::UploadTexelData { .. } => simple(ffi::UPLOAD_TEXEL_DATA),
            RenderCommand::DeclareRenderTarget { .. } => simple(ffi::DECLARE_RENDER_TARGET),
        };
        check(unsafe { ffi::PFCudaRendererRenderCommand(self.raw, &c) })
    }

    /// RGBA8 rows, top-left origin, into `dst` (`stride` bytes per row).
    pub fn read_pixels(&mut self, dst: &mut [u8], stride: usize) {
        check(unsafe { ffi::PFCudaRendererReadPixels(self.raw, dst.as_mut_ptr(), stride) })
    }
}

impl Drop for CudaRenderer {
    fn drop(&mut self) {
        unsafe { ffi::PFCudaRendererDestroy(self.raw) }
    }
}

/// `Scene::build_and_render` (renderer/src/scene.rs:369-378) for the CUDA backend.
pub fn build_and_render<E: Executor>(scene: &mut Scene, renderer: &mut CudaRenderer, options: BuildOptions,
                                     executor: E) {
    // The command stream does not carry scene.view_box(), which process_line_segment clips to
    // (renderer/src/tiler.rs:194).
    renderer.set_view_box(scene.view_box());
    let commands = Arc::new(Mutex::new(vec![]));
    let sink_commands = commands.clone();
    let listener = RenderCommandListener::new(Box::new(move |command| sink_commands.lock().unwrap().push(command)));
    let mut sink = SceneSink::new(listener, RendererLevel::D3D11);
    scene.build(options, &mut sink, &executor);
    renderer.begin_scene();
    for command in commands.lock().unwrap().iter() {
        renderer.render_command(command);
    }
    renderer.end_scene();
}
