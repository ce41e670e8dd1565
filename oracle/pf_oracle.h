/* oracle/pf_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU oracle: a plain C++ restatement of Pathfinder 3's CPU
 * tiler (D3D9 level, SequentialExecutor order) and of the reference's fill /
 * composite shader math. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library. The product
 * (pathfinder_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference holds no tests or golden vectors for this path
 *  and cannot be compiled here (no rustc/cargo); this restatement
 * is pinned only by the weak simd KATs (simd/src/test.rs), by its own property
 * tests (tests/test_oracle.py), by the reference's shipped LUT textures, and by
 * double entry: tests/py_tiler.py restates the tiler, the batch pack and the
 * shader math a second time, independently, and tests/test_py_tiler.py demands
 * bit-identical lists (float-tolerant pixels) from both. Neither is the Rust
 * tiler itself: tools/reference_dump/ is the kit that pins it where cargo exists.
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Point flags, as content/src/outline.rs PointFlags (CONTROL_POINT_0 = 1, CONTROL_POINT_1 = 2). */
#define PFO_POINT_ON_CURVE 0u
#define PFO_POINT_CTRL0 1u
#define PFO_POINT_CTRL1 2u

#define PFO_FILL_RULE_WINDING 0u
#define PFO_FILL_RULE_EVEN_ODD 1u

#define PFO_NO_CLIP 0xffffffffu

/* A scene as flat arrays (all borrowed). Paths own contiguous contour ranges;
 * contours own contiguous point ranges. Clip paths and draw paths share the
 * contour/point pools. */
typedef struct PFOScene {
    const float *points;               /* [n_points][2] */
    const uint8_t *point_flags;        /* [n_points] */
    const uint32_t *contour_offsets;   /* [n_contours + 1] -> point index */
    uint32_t n_points;
    uint32_t n_contours;

    uint32_t n_clip_paths;
    const uint32_t *clip_contour_ranges; /* [n_clip_paths][2] first contour, end contour */
    const uint8_t *clip_fill_rules;      /* [n_clip_paths] */

    uint32_t n_draw_paths;
    const uint32_t *draw_contour_ranges; /* [n_draw_paths][2] */
    const uint8_t *draw_fill_rules;      /* [n_draw_paths] */
    const uint16_t *draw_paints;         /* [n_draw_paths] paint id */
    const uint32_t *draw_clip_paths;     /* [n_draw_paths] clip path id or PFO_NO_CLIP */

    uint32_t n_paints;
    const uint8_t *paint_colors;         /* [n_paints][4] RGBA8, solid colours only */

    float view_box[4];                   /* min_x, min_y, max_x, max_y */
} PFOScene;

/* BuildOptions (renderer/src/options.rs:53-61), 2-D transforms only. */
typedef struct PFOBuildOptions {
    int32_t has_transform;   /* 0 = RenderTransform::default() (identity 2D) */
    float transform[6];      /* m11, m21, m12, m22, tx, ty  (geometry/src/transform2d.rs) */
    float dilation[2];
    int32_t subpixel_aa_enabled;
    /* Strip restriction for the multi-GPU partition check: tile rows [y0, y1).
     * y0 = y1 = 0 means "no restriction". Not a reference feature: it restates what a
     * rank that owns those tile rows must produce. */
    int32_t strip_tile_y0, strip_tile_y1;
} PFOBuildOptions;

/* Mirrors of the reference's #[repr(C)] records (renderer/src/gpu_data.rs). */
typedef struct PFOFill {          /* gpu_data.rs:354-363 */
    uint16_t from_x, from_y, to_x, to_y;
    uint32_t link;                /* D3D9: alpha tile id */
} PFOFill;

typedef struct PFOTileObjectPrimitive { /* gpu_data.rs:264-275 */
    int16_t tile_x, tile_y;
    uint32_t alpha_tile_id;
    uint32_t path_id;
    uint16_t color;
    uint8_t ctrl;
    int8_t backdrop;
} PFOTileObjectPrimitive;

typedef struct PFOClip {          /* gpu_data.rs:376-383 */
    uint32_t dest_tile_id;
    int32_t dest_backdrop;
    uint32_t src_tile_id;
    int32_t src_backdrop;
} PFOClip;

typedef struct PFOBuilt PFOBuilt;

/* Runs the CPU tiler over the whole scene in SequentialExecutor order (clip paths 0..n, then
 * draw paths 0..m). n_threads > 1 uses std::thread workers pulling path indices from an atomic
 * counter (the Rayon par_iter stand-in); alpha tile ids are then racy exactly as in the
 * reference, so only n_threads == 1 results are canonical. */
PFOBuilt *pfo_build(const PFOScene *scene, const PFOBuildOptions *options, int n_threads);
void pfo_built_destroy(PFOBuilt *built);

/* Accessors. Pointers stay valid until pfo_built_destroy. */
size_t pfo_fill_count(const PFOBuilt *b);
const PFOFill *pfo_fills(const PFOBuilt *b);             /* all AddFillsD3D9 payloads, in path order */
const uint32_t *pfo_fill_path_offsets(const PFOBuilt *b); /* [n_clip+n_draw+1] into fills */
size_t pfo_tile_count(const PFOBuilt *b);
const PFOTileObjectPrimitive *pfo_tiles(const PFOBuilt *b); /* DrawTileBatchD3D9.tiles */
size_t pfo_clip_count(const PFOBuilt *b);
const PFOClip *pfo_clips(const PFOBuilt *b);
const int32_t *pfo_z_buffer(const PFOBuilt *b, int32_t rect_out[4]); /* DenseTileMap<i32> */
uint32_t pfo_alpha_tile_count(const PFOBuilt *b);
uint64_t pfo_line_segment_count(const PFOBuilt *b);  /* process_line_segment calls ("segments") */
uint64_t pfo_input_segment_count(const PFOBuilt *b); /* contour iterator items */
uint64_t pfo_bbox_tile_count(const PFOBuilt *b);     /* sum of dense tile map areas */
double pfo_build_seconds(const PFOBuilt *b);         /* wall time of the build (cpu_build_time) */
double pfo_build_seconds_paths(const PFOBuilt *b);   /* its per-path (parallel) part */

/* Flattened line segments of one path in emission order (for dice parity). Returns count;
 * copies up to cap segments (4 floats each) into out. path index is in [0, n_clip + n_draw). */
size_t pfo_path_lines(const PFOBuilt *b, uint32_t path, float *out, size_t cap);
/* Keep flattened lines (costs memory); off by default. Call before pfo_build. */
void pfo_set_keep_lines(int keep);

/* Coverage mask of every alpha tile: out[alpha_tile_id][16 rows][16 cols] f32 = sum of
 * computeCoverage over the tile's fills (shaders/fill_area.inc.glsl:11-27, evaluated in 4-row
 * strips exactly as shaders/d3d9/fill.vs.glsl:46-66 / d3d11/fill_compute.inc.glsl:11-25),
 * without backdrop, unclamped (D3D9 RGBA16F mask semantics, kept in f32). area_lut is the
 * 256x256 RGBA8 texture (bilinear, clamp to edge). */
void pfo_alpha_masks(const PFOBuilt *b, const uint8_t *area_lut_rgba, float *out);

/* Composites the batch into an RGBA8 image of width x height pixels (row-major, top-left
 * origin), following shaders/d3d11/tile.cs.glsl:71-163 with mask semantics of
 * shaders/tile_fragment.inc.glsl:539-556 on the f32 mask. background = clear colour (RGBA f32,
 * premultiplied as given). out_f32 (optional, may be NULL) receives the unrounded floats. */
void pfo_render(const PFOBuilt *b, const PFOScene *scene, const uint8_t *area_lut_rgba,
                const float background[4], uint32_t width, uint32_t height, uint8_t *out_rgba,
                float *out_f32);

/* The same for the pixels [x0, x0 + width) x [y0, y0 + height) of a frame_w x frame_h frame (out_rgba /
 * out_f32 hold width x height pixels). Only the masks of alpha tiles that reach the crop are evaluated, so
 * crops of frames too large to composite whole on the CPU stay cheap. */
void pfo_render_crop(const PFOBuilt *b, const PFOScene *scene, const uint8_t *area_lut_rgba,
                     const float background[4], uint32_t frame_w, uint32_t frame_h, uint32_t x0, uint32_t y0,
                     uint32_t width, uint32_t height, uint8_t *out_rgba, float *out_f32);

#ifdef __cplusplus
}
#endif
#endif
