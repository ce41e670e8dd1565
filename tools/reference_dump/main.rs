// tools/reference_dump/main.rs — prints the reference's own tile / fill / z-buffer lists for a scene
// file written by `tools/dump_lists.py --scene-out`, in the format tools/diff_lists.py reads.
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE (no Rust toolchain): this is the program a
// maintainer with cargo runs once to pin the oracle against the real tiler. Inside a checkout of
// servo/pathfinder:
//
//     mkdir -p examples/reference_dump/src && cp main.rs examples/reference_dump/src/
//     # Cargo.toml of the example: dependencies pathfinder_renderer (features = ["d3d9"]),
//     #   pathfinder_content, pathfinder_geometry, pathfinder_color (path = "../../<crate>"); add the
//     #   example to the workspace members. `gpu_data` is a private module of pathfinder_renderer
//     #   (renderer/src/lib.rs:28): make it `pub mod gpu_data;` for this build.
//     cargo run --release -p reference_dump -- tiger1024.scene > reference.lists
//     python tools/diff_lists.py ours.lists reference.lists
//
// It builds the Scene from the outlines in the file (the comparison origin: post-loader outlines), runs
// `Scene::build` at RendererLevel::D3D9 with SequentialExecutor (the canonical order of SURVEY.md §8c)
// and prints every AddFillsD3D9 / DrawTilesD3D9 payload, Clip records included (renderer/src/gpu_data.rs:69,97,
// 227-237,266-275,356-363,378-383). Clip paths are read too (`clippath` records, `path ... <clip id>`).

use pathfinder_color::ColorU;
use pathfinder_content::fill::FillRule;
use pathfinder_content::outline::{Contour, Outline};
use pathfinder_geometry::rect::RectF;
use pathfinder_geometry::transform2d::{Matrix2x2F, Transform2F};
use pathfinder_geometry::vector::{vec2f, Vector2F};
use pathfinder_renderer::concurrent::executor::SequentialExecutor;
use pathfinder_renderer::gpu::options::RendererLevel;
use pathfinder_renderer::gpu_data::RenderCommand;
use pathfinder_renderer::options::{BuildOptions, RenderCommandListener, RenderTransform};
use pathfinder_renderer::paint::Paint;
use pathfinder_renderer::scene::{ClipPath, ClipPathId, DrawPath, Scene, SceneSink};
use std::env;
use std::fs;
use std::sync::{Arc, Mutex};

fn f(bits: &str) -> f32 {
    f32::from_bits(u32::from_str_radix(bits, 16).unwrap())
}

fn main() {
    let path = env::args().nth(1).expect("usage: reference_dump <scene file>");
    let text = fs::read_to_string(path).unwrap();
    let mut scene = Scene::new();
    let mut transform = Transform2F::default();
    let mut dilation = vec2f(0.0, 0.0);
    let mut paints = vec![];
    // (fill rule, paint (None: a clip path), clip path of a draw path, outline)
    let mut current: Option<(FillRule, Option<usize>, Option<ClipPathId>, Outline)> = None;
    let mut clip_ids: Vec<ClipPathId> = vec![];
    let mut contour: Option<(Contour, Vec<(Vector2F, u8)>, usize)> = None;

    fn finish_contour(points: &[(Vector2F, u8)]) -> Contour {
        // Points carry the reference's own flags (content/src/outline.rs PointFlags): an on-curve point
        // preceded by 0, 1 or 2 control points is a line, quadratic or cubic end point.
        let mut c = Contour::new();
        let mut i = 0;
        while i < points.len() {
            let (p, flags) = points[i];
            if flags == 0 {
                c.push_endpoint(p);
                i += 1;
            } else if i + 1 < points.len() && points[i + 1].1 == 0 {
                c.push_quadratic(p, points[i + 1].0);
                i += 2;
            } else {
                c.push_cubic(p, points[i + 1].0, points[i + 2].0);
                i += 3;
            }
        }
        c.close();
        c
    }

    let mut flush_path = |scene: &mut Scene, paints: &Vec<_>, clip_ids: &mut Vec<ClipPathId>,
                          cur: (FillRule, Option<usize>, Option<ClipPathId>, Outline)| {
        match cur.1 {
            Some(paint) => {
                let mut draw_path = DrawPath::new(cur.3, paints[paint]);
                draw_path.set_fill_rule(cur.0);
                draw_path.set_clip_path(cur.2);
                scene.push_draw_path(draw_path);
            }
            None => {
                let mut clip_path = ClipPath::new(cur.3);
                clip_path.set_fill_rule(cur.0);
                clip_ids.push(scene.push_clip_path(clip_path));
            }
        }
    };

    for line in text.lines() {
        let w: Vec<&str> = line.split_whitespace().collect();
        match w.get(0).copied() {
            Some("viewbox") => scene.set_view_box(RectF::from_points(vec2f(f(w[1]), f(w[2])), vec2f(f(w[3]), f(w[4])))),
            Some("transform") => {
                transform = Transform2F {
                    matrix: Matrix2x2F::row_major(f(w[1]), f(w[2]), f(w[3]), f(w[4])),
                    vector: vec2f(f(w[5]), f(w[6])),
                }
            }
            Some("dilation") => dilation = vec2f(f(w[1]), f(w[2])),
            Some("paint") => {
                let c: Vec<u8> = w[1..5].iter().map(|v| v.parse().unwrap()).collect();
                paints.push(scene.push_paint(&Paint::from_color(ColorU::new(c[0], c[1], c[2], c[3]))));
            }
            Some("clippath") => {
                if let Some(cur) = current.take() {
                    flush_path(&mut scene, &paints, &mut clip_ids, cur);
                }
                let rule = if w[1] == "1" { FillRule::EvenOdd } else { FillRule::Winding };
                current = Some((rule, None, None, Outline::new()));
            }
            Some("path") => {
                if let Some(cur) = current.take() {
                    flush_path(&mut scene, &paints, &mut clip_ids, cur);
                }
                let rule = if w[1] == "1" { FillRule::EvenOdd } else { FillRule::Winding };
                let clip: i64 = w.get(4).map_or(-1, |v| v.parse().unwrap());
                let clip = if clip < 0 { None } else { Some(clip_ids[clip as usize]) };
                current = Some((rule, Some(w[2].parse().unwrap()), clip, Outline::new()));
            }
            Some("contour") => contour = Some((Contour::new(), vec![], w[1].parse().unwrap())),
            Some("p") => {
                let done = {
                    let c = contour.as_mut().unwrap();
                    c.1.push((vec2f(f(w[1]), f(w[2])), w[3].parse().unwrap()));
                    c.1.len() == c.2
                };
                if done {
                    let c = contour.take().unwrap();
                    current.as_mut().unwrap().3.push_contour(finish_contour(&c.1));
                }
            }
            _ => {}
        }
    }
    if let Some(cur) = current.take() {
        flush_path(&mut scene, &paints, &mut clip_ids, cur);
    }

    let commands = Arc::new(Mutex::new(vec![]));
    let sink_commands = commands.clone();
    let listener = RenderCommandListener::new(Box::new(move |command| sink_commands.lock().unwrap().push(command)));
    let mut sink = SceneSink::new(listener, RendererLevel::D3D9);
    let options = BuildOptions {
        transform: RenderTransform::Transform2D(transform),
        dilation,
        ..BuildOptions::default()
    };
    scene.build(options, &mut sink, &SequentialExecutor);

    for command in commands.lock().unwrap().iter() {
        match *command {
            RenderCommand::AddFillsD3D9(ref fills) => {
                for fill in fills {
                    let s = fill.line_segment;
                    println!("fill {} {} {} {} {}", s.from_x, s.from_y, s.to_x, s.to_y, fill.link);
                }
            }
            RenderCommand::DrawTilesD3D9(ref batch) => {
                for t in &batch.tiles {
                    println!("tile {} {} {} {} {} {} {}", t.tile_x, t.tile_y, t.alpha_tile_id.0, t.path_id.0,
                             t.color, t.ctrl, t.backdrop);
                }
                let z = &batch.z_buffer_data;
                let texels: Vec<String> = z.data.iter().map(|v| v.to_string()).collect();
                println!("z {} {} {}", z.rect.width(), z.rect.height(), texels.join(" "));
                for c in &batch.clips {
                    println!("clip {} {} {} {}", c.dest_tile_id.0, c.dest_backdrop, c.src_tile_id.0, c.src_backdrop);
                }
            }
            _ => {}
        }
    }
}
