//! A CUDA backend for Pathfinder 3's D3D11-level pipeline. It plugs in at the renderer protocol
//! (`begin_scene` / `render_command` / `end_scene`, renderer/src/gpu/renderer.rs:350-460) rather than
//! at `trait Device`, which is shaped around GLSL programs and raster draws.
//!
//! One change to the reference is needed to build this outside `pathfinder_renderer`: its `gpu_data`
//! module (RenderCommand and the batch records) is private (`renderer/src/lib.rs:28`); make it
//! `pub mod gpu_data;`, or move this file in next to `gpu/d3d11/renderer.rs` and `crate::` the paths.

mod ffi;

use pathfinder_color::ColorF;
use pathfinder_content::effects::{BlendMode, BlurDirection, Filter, PatternFilter};
use pathfinder_geometry::rect::RectF;
use pathfinder_geometry::vector::Vector2I;
use pathfinder_renderer::concurrent::executor::Executor;
use pathfinder_renderer::gpu::options::{RendererLevel, RendererMode};
use pathfinder_renderer::gpu_data::{ColorCombineMode, PathSource, RenderCommand, SegmentsD3D11};
use pathfinder_renderer::gpu_data::{TextureLocation, TextureMetadataEntry, TileBatchDataD3D11, TileBatchTexture};
use pathfinder_renderer::options::{BuildOptions, RenderCommandListener};
use pathfinder_renderer::scene::{Scene, SceneSink};
use pathfinder_resources::ResourceLoader;
use std::ffi::CStr;
use std::os::raw::c_void;
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

// BlendMode -> PF_BLEND_MODE_*: an explicit table, not `as u32` (the header does not promise to follow the
// order of the Rust enum, content/src/effects.rs:99-163).
fn blend_mode(mode: BlendMode) -> u32 {
    match mode {
        BlendMode::Clear => ffi::BLEND_MODE_CLEAR,
        BlendMode::Copy => ffi::BLEND_MODE_COPY,
        BlendMode::SrcIn => ffi::BLEND_MODE_SRC_IN,
        BlendMode::SrcOut => ffi::BLEND_MODE_SRC_OUT,
        BlendMode::SrcOver => ffi::BLEND_MODE_SRC_OVER,
        BlendMode::SrcAtop => ffi::BLEND_MODE_SRC_ATOP,
        BlendMode::DestIn => ffi::BLEND_MODE_DEST_IN,
        BlendMode::DestOut => ffi::BLEND_MODE_DEST_OUT,
        BlendMode::DestOver => ffi::BLEND_MODE_DEST_OVER,
        BlendMode::DestAtop => ffi::BLEND_MODE_DEST_ATOP,
        BlendMode::Xor => ffi::BLEND_MODE_XOR,
        BlendMode::Lighter => ffi::BLEND_MODE_LIGHTER,
        BlendMode::Darken => ffi::BLEND_MODE_DARKEN,
        BlendMode::Lighten => ffi::BLEND_MODE_LIGHTEN,
        BlendMode::Multiply => ffi::BLEND_MODE_MULTIPLY,
        BlendMode::Screen => ffi::BLEND_MODE_SCREEN,
        BlendMode::HardLight => ffi::BLEND_MODE_HARD_LIGHT,
        BlendMode::Overlay => ffi::BLEND_MODE_OVERLAY,
        BlendMode::ColorDodge => ffi::BLEND_MODE_COLOR_DODGE,
        BlendMode::ColorBurn => ffi::BLEND_MODE_COLOR_BURN,
        BlendMode::SoftLight => ffi::BLEND_MODE_SOFT_LIGHT,
        BlendMode::Difference => ffi::BLEND_MODE_DIFFERENCE,
        BlendMode::Exclusion => ffi::BLEND_MODE_EXCLUSION,
        BlendMode::Hue => ffi::BLEND_MODE_HUE,
        BlendMode::Saturation => ffi::BLEND_MODE_SATURATION,
        BlendMode::Color => ffi::BLEND_MODE_COLOR,
        BlendMode::Luminosity => ffi::BLEND_MODE_LUMINOSITY,
    }
}

// Filter (content/src/effects.rs:44-97) -> PFFilter: kind, flags and up to 20 floats, in the order pf_cuda.h states.
fn filter(f: &Filter) -> ffi::PFFilter {
    let mut out = ffi::PFFilter { kind: ffi::FILTER_NONE, flags: 0, params: [0.0; 20] };
    match *f {
        Filter::None => {}
        Filter::RadialGradient { line, radii, uv_origin } => {
            out.kind = ffi::FILTER_RADIAL_GRADIENT;
            out.params[..8].copy_from_slice(&[line.from_x(), line.from_y(), line.to_x(), line.to_y(),
                                              radii.x(), radii.y(), uv_origin.x(), uv_origin.y()]);
        }
        Filter::PatternFilter(PatternFilter::Text { fg_color, bg_color, defringing_kernel, gamma_correction }) => {
            out.kind = ffi::FILTER_TEXT;
            out.params[..8].copy_from_slice(&[fg_color.r(), fg_color.g(), fg_color.b(), fg_color.a(),
                                              bg_color.r(), bg_color.g(), bg_color.b(), bg_color.a()]);
            if let Some(kernel) = defringing_kernel {
                out.flags |= ffi::FILTER_FLAG_TEXT_HAS_KERNEL;
                out.params[8..12].copy_from_slice(&kernel.0);
            }
            if gamma_correction {
                out.flags |= ffi::FILTER_FLAG_TEXT_GAMMA_CORRECTION;
            }
        }
        Filter::PatternFilter(PatternFilter::Blur { direction, sigma }) => {
            out.kind = ffi::FILTER_BLUR;
            out.params[0] = sigma;
            if let BlurDirection::Y = direction {
                out.flags |= ffi::FILTER_FLAG_BLUR_Y;
            }
        }
        Filter::PatternFilter(PatternFilter::ColorMatrix(ref matrix)) => {
            out.kind = ffi::FILTER_COLOR_MATRIX;
            for (column, lanes) in matrix.0.iter().enumerate() {
                out.params[column * 4..column * 4 + 4]
                    .copy_from_slice(&[lanes.x(), lanes.y(), lanes.z(), lanes.w()]);
            }
        }
    }
    out
}

// TextureMetadataEntry (renderer/src/gpu_data.rs:336-344) -> PFTextureMetadataEntry, field by field. The
// transform is read through its accessors, as Renderer::upload_texture_metadata does (gpu/renderer.rs:712-722),
// so the lane order of the F32x4 behind Matrix2x2F never matters here.
fn texture_metadata_entry(e: &TextureMetadataEntry) -> ffi::PFTextureMetadataEntry {
    let t = &e.color_0_transform;
    ffi::PFTextureMetadataEntry {
        color_0_transform: [t.m11(), t.m12(), t.m21(), t.m22(), t.m13(), t.m23()],
        color_0_combine_mode: match e.color_0_combine_mode {
            ColorCombineMode::None => ffi::COLOR_COMBINE_MODE_NONE,
            ColorCombineMode::SrcIn => ffi::COLOR_COMBINE_MODE_SRC_IN,
            ColorCombineMode::DestIn => ffi::COLOR_COMBINE_MODE_DEST_IN,
        },
        base_color: [e.base_color.r, e.base_color.g, e.base_color.b, e.base_color.a],
        blend_mode: blend_mode(e.blend_mode),
        filter: filter(&e.filter),
    }
}

fn texture_location(l: &TextureLocation) -> ffi::PFTextureLocation {
    // RectI is SIMD lanes with 16-byte alignment: copied out lane by lane, never cast.
    ffi::PFTextureLocation {
        page: l.page.0,
        rect: [l.rect.origin_x(), l.rect.origin_y(), l.rect.lower_right().x(), l.rect.lower_right().y()],
    }
}

fn tile_batch_texture(t: &Option<TileBatchTexture>) -> ffi::PFTileBatchTexture {
    match *t {
        None => ffi::PFTileBatchTexture { page: 0, sampling_flags: 0, composite_op: ffi::PAINT_COMPOSITE_OP_SRC_IN },
        Some(ref t) => ffi::PFTileBatchTexture {
            page: t.page.0,
            sampling_flags: t.sampling_flags.bits(), // the PF_TEXTURE_SAMPLING_FLAGS_* bits are the reference's
            // `composite_op` is pub(crate) in the reference (gpu_data.rs:252). Inside the crate (this file moved
            // next to gpu/d3d11/renderer.rs) read the field; outside, either add a one-line accessor there or,
            // as here, tell the two variants apart through the derived Debug.
            composite_op: if format!("{:?}", t).contains("DestIn") { ffi::PAINT_COMPOSITE_OP_DEST_IN }
                          else { ffi::PAINT_COMPOSITE_OP_SRC_IN },
        },
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
        let area = image::load_from_memory(&resources.slurp("textures/area-lut.png").unwrap()).unwrap().to_rgba8();
        // textures/gamma-lut.png, the 256 x 8 L8 table of the text filter (gpu/renderer.rs:207-222 loads both)
        let gamma = image::load_from_memory(&resources.slurp("textures/gamma-lut.png").unwrap()).unwrap().to_luma8();
        let options = ffi::PFCudaRendererOptions {
            dest_size: [dest_size.x(), dest_size.y()],
            background_color: background.map_or([0.0; 4], |c| [c.r(), c.g(), c.b(), c.a()]),
            flags: background.is_some() as u8,
        };
        let raw = unsafe {
            ffi::PFCudaRendererCreate(ffi::PFCudaDeviceCreate(ordinal), area.as_ptr(), gamma.as_ptr(),
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
        // (kept alive until the call returns: the command borrows it)
        let converted_metadata: Vec<ffi::PFTextureMetadataEntry>;
        let c = match *command {
            RenderCommand::Start { path_count, needs_readable_framebuffer, .. } => ffi::PFRenderCommand {
                kind: ffi::START,
                u: ffi::PFRenderCommandPayload {
                    start: ffi::PFStart { path_count: path_count as u64,
                                          needs_readable_framebuffer: needs_readable_framebuffer as u32 },
                },
            },
            RenderCommand::UploadTextureMetadata(ref entries) => {
                converted_metadata = entries.iter().map(texture_metadata_entry).collect();
                ffi::PFRenderCommand {
                    kind: ffi::UPLOAD_TEXTURE_METADATA,
                    u: ffi::PFRenderCommandPayload {
                        upload_texture_metadata: ffi::PFUploadTextureMetadata {
                            entries: converted_metadata.as_ptr(),
                            entry_count: converted_metadata.len(),
                            content_key: 0,
                        },
                    },
                }
            }
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
                        color_texture: tile_batch_texture(&draw.color_texture),
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
            RenderCommand::AllocateTexturePage { page_id, ref descriptor } => ffi::PFRenderCommand {
                kind: ffi::ALLOCATE_TEXTURE_PAGE,
                u: ffi::PFRenderCommandPayload {
                    allocate_texture_page: ffi::PFAllocateTexturePage {
                        page_id: page_id.0,
                        size: [descriptor.size.x(), descriptor.size.y()],
                    },
                },
            },
            RenderCommand::UploadTexelData { ref texels, ref location } => ffi::PFRenderCommand {
                kind: ffi::UPLOAD_TEXEL_DATA,
                u: ffi::PFRenderCommandPayload {
                    upload_texel_data: ffi::PFUploadTexelData {
                        texels: texels.as_ptr(),
                        texel_count: texels.len(),
                        location: texture_location(location),
                    },
                },
            },
            RenderCommand::DeclareRenderTarget { id, ref location } => ffi::PFRenderCommand {
                kind: ffi::DECLARE_RENDER_TARGET,
                u: ffi::PFRenderCommandPayload {
                    declare_render_target: ffi::PFDeclareRenderTarget {
                        render_target_id: id.render_target,
                        location: texture_location(location),
                    },
                },
            },
            // D3D9-level commands are refused with PF_CUDA_ERROR_WRONG_LEVEL, paints / blends / filters outside
            // the built set with PF_CUDA_ERROR_UNSUPPORTED: `check` turns both into a panic, like the reference.
            RenderCommand::AddFillsD3D9(_) => simple(ffi::ADD_FILLS_D3D9),
            RenderCommand::FlushFillsD3D9 => simple(ffi::FLUSH_FILLS_D3D9),
            RenderCommand::DrawTilesD3D9(_) => simple(ffi::DRAW_TILES_D3D9),
        };
        check(unsafe { ffi::PFCudaRendererRenderCommand(self.raw, &c) })
    }

    /// RGBA8 rows, top-left origin, into `dst` (`stride` bytes per row).
    /// Multi-GPU frame assembly (one process per GPU; include/pf_cuda.h "PFCudaRendererGather*"). `id` comes from
    /// `gather_create_id()` on one rank and reaches the others through any host channel.
    pub fn gather_create_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        check(unsafe { ffi::PFCudaGatherCreateId(id.as_mut_ptr()) });
        id
    }

    /// Collective: joins the group and sets this renderer's strip of tile rows.
    pub fn gather_init(&mut self, id: &[u8; 128], rank: i32, world_size: i32) {
        check(unsafe { ffi::PFCudaRendererGatherInit(self.raw, id.as_ptr(), rank, world_size) })
    }

    /// 0 = all-gather of the finished strips, 1 = compact tile exports pushed over NVLink (the default where possible).
    pub fn gather_set_mode(&mut self, mode: i32) {
        check(unsafe { ffi::PFCudaRendererGatherSetMode(self.raw, mode) })
    }

    /// Collective, asynchronous: completes every rank's copy of the frame just rendered.
    pub fn gather_frame(&mut self) {
        check(unsafe { ffi::PFCudaRendererGatherFrame(self.raw) })
    }

    pub fn gather_wait(&mut self) {
        check(unsafe { ffi::PFCudaRendererGatherWait(self.raw) })
    }

    pub fn gather_destroy(&mut self) {
        check(unsafe { ffi::PFCudaRendererGatherDestroy(self.raw) })
    }

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

/// `SceneProxy` (renderer/src/concurrent/scene_proxy.rs:35-157) for the CUDA backend. The reference's own proxy cannot
/// be reused as it is: its `render` / `build_and_render` take `&mut Renderer<D>` and its receiver is private. This is the
/// same thing — the scene on a thread of its own, messages in, owned `RenderCommand`s back through a channel — with
/// `CudaRenderer` at the receiving end. (The C library has the equivalent for hosts without the Rust scene:
/// `PFSceneProxyCreateFromScene` .. `PFSceneProxyBuildAndRenderCuda`, include/pf_cuda.h.)
pub struct CudaSceneProxy {
    sender: std::sync::mpsc::SyncSender<ProxyMsg>,
    receiver: std::sync::mpsc::Receiver<RenderCommand>,
    view_boxes: std::collections::VecDeque<RectF>, // the view box each queued build will see (the stream does not carry it)
    view_box: RectF,
}

enum ProxyMsg {
    ReplaceScene(Scene),
    CopyScene(std::sync::mpsc::SyncSender<Scene>),
    SetViewBox(RectF),
    Build(BuildOptions),
}

const MAX_MESSAGES_IN_FLIGHT: usize = 1024; // scene_proxy.rs:33

impl CudaSceneProxy {
    /// `SceneProxy::from_scene` (scene_proxy.rs:60-74), always at `RendererLevel::D3D11`.
    pub fn from_scene<E: Executor + Send + 'static>(scene: Scene, executor: E) -> CudaSceneProxy {
        let (to_worker, from_main) = std::sync::mpsc::sync_channel(MAX_MESSAGES_IN_FLIGHT);
        let (to_main, from_worker) = std::sync::mpsc::sync_channel(MAX_MESSAGES_IN_FLIGHT);
        let to_main = Mutex::new(to_main);
        let listener = RenderCommandListener::new(Box::new(move |command| drop(to_main.lock().unwrap().send(command))));
        let view_box = scene.view_box();
        std::thread::spawn(move || {
            let mut scene = scene;
            let mut sink = SceneSink::new(listener, RendererLevel::D3D11);
            while let Ok(msg) = from_main.recv() {
                match msg {
                    ProxyMsg::ReplaceScene(new_scene) => scene = new_scene,
                    ProxyMsg::CopyScene(reply) => drop(reply.send(scene.clone())),
                    ProxyMsg::SetViewBox(view_box) => scene.set_view_box(view_box),
                    ProxyMsg::Build(options) => scene.build(options, &mut sink, &executor),
                }
            }
        });
        CudaSceneProxy { sender: to_worker, receiver: from_worker, view_boxes: Default::default(), view_box }
    }

    pub fn replace_scene(&mut self, new_scene: Scene) {
        self.view_box = new_scene.view_box();
        self.sender.send(ProxyMsg::ReplaceScene(new_scene)).unwrap();
    }

    pub fn set_view_box(&mut self, new_view_box: RectF) {
        self.view_box = new_view_box;
        self.sender.send(ProxyMsg::SetViewBox(new_view_box)).unwrap();
    }

    pub fn build(&mut self, options: BuildOptions) {
        self.view_boxes.push_back(self.view_box);
        self.sender.send(ProxyMsg::Build(options)).unwrap();
    }

    /// `SceneProxy::render` (scene_proxy.rs:95-105).
    pub fn render(&mut self, renderer: &mut CudaRenderer) {
        renderer.set_view_box(self.view_boxes.pop_front().expect("render without a build"));
        renderer.begin_scene();
        while let Ok(command) = self.receiver.recv() {
            renderer.render_command(&command);
            if let RenderCommand::Finish { .. } = command {
                break;
            }
        }
        renderer.end_scene();
    }

    pub fn build_and_render(&mut self, renderer: &mut CudaRenderer, options: BuildOptions) {
        self.build(options);
        self.render(renderer);
    }

    pub fn copy_scene(&self) -> Scene {
        let (reply, scene) = std::sync::mpsc::sync_channel(1);
        self.sender.send(ProxyMsg::CopyScene(reply)).unwrap();
        scene.recv().unwrap()
    }
}
