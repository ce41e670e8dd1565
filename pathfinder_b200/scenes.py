"""Benchmark / test scenes (BASELINE.json configs).

* tiger: Ghostscript tiger outlines, loaded from the committed fixture tests/golden/tiger.npz
  (produced by tools/make_tiger_scene.py from resources/svg/Ghostscript_Tiger.svg).
* random_paths: the synthetic "N random overlapping cubic-Bézier paths" scenes of configs 4 and 5.
"""
from __future__ import annotations

import os

import numpy as np

from .flat_scene import FILL_RULE_EVEN_ODD, FILL_RULE_WINDING, FlatScene

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIGER_FIXTURE = os.path.join(_ROOT, "tests", "golden", "tiger.npz")

_M64 = (1 << 64) - 1


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """Outputs start .. start+count-1 of the SplitMix64 stream seeded with `seed` (vectorised: the
    state after i steps is seed + i * 0x9E3779B97F4A7C15)."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed & _M64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _unit(u64: np.ndarray) -> np.ndarray:
    """u64 -> float64 in [0, 1) from the top 53 bits."""
    return (u64 >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


DRAWS_PER_PATH = 64
PALETTE_SIZE = 4096  # PaintId is u16 and the metadata texture holds 65,536 entries (paint.rs:81)


def random_paths(n_paths: int, size: int, seed: int, r_min: float = 16.0, r_max: float = 256.0,
                 name: str | None = None) -> FlatScene:
    """Config 4/5 generator: per path one closed contour of k in {4..8} cubic
    segments around a random centre; radius log-uniform in [r_min, r_max]; end points at angle
    2*pi*m/k and radius r*U(0.6, 1.0); control points = end points +/- tangent * r * U(0.2, 0.6);
    fill rule alternates by index; colour from a seeded 4096-entry palette whose entries are
    opaque with probability 1/2, else alpha in 64..254.

    Path i reads draws [i*64, i*64+64) of SplitMix64(seed); the palette reads the draws after the
    last path, so a scene with fewer paths is a prefix of a larger one (same seed)."""
    u = _unit(splitmix64(seed, 0, n_paths * DRAWS_PER_PATH)).reshape(n_paths, DRAWS_PER_PATH)
    cx, cy = u[:, 0] * size, u[:, 1] * size
    r = r_min * (r_max / r_min) ** u[:, 2]
    k = 4 + np.minimum((u[:, 3] * 5).astype(np.int64), 4)
    pal_idx = np.minimum((u[:, 4] * PALETTE_SIZE).astype(np.int64), PALETTE_SIZE - 1)

    kmax = 8
    m = np.arange(kmax)
    theta = 2.0 * np.pi * m[None, :] / k[:, None]                       # (n, 8)
    rad = r[:, None] * (0.6 + 0.4 * u[:, 8:16])
    ex, ey = cx[:, None] + rad * np.cos(theta), cy[:, None] + rad * np.sin(theta)
    tx, ty = -np.sin(theta), np.cos(theta)
    out_len = r[:, None] * (0.2 + 0.4 * u[:, 16:24])                    # leaving the end point
    in_len = r[:, None] * (0.2 + 0.4 * u[:, 24:32])                     # arriving at the next one
    nxt = (m[None, :] + 1) % k[:, None]
    rows = np.arange(n_paths)[:, None]
    c0x, c0y = ex + tx * out_len, ey + ty * out_len
    c1x = ex[rows, nxt] - tx[rows, nxt] * in_len
    c1y = ey[rows, nxt] - ty[rows, nxt] * in_len

    # Contour layout (content/src/outline.rs): P0, then (c0, c1, P_{m+1}) for m = 0..k-2, then
    # (c0, c1) of the last segment followed by... the closing point. The last cubic must end on an
    # on-curve point, so it ends at a copy of P0; the implicit closing line is then zero-length.
    pts_per_path = 1 + 3 * k
    point_offsets = np.concatenate([[0], np.cumsum(pts_per_path)]).astype(np.int64)
    n_points = int(point_offsets[-1])
    points = np.zeros((n_points, 2), dtype=np.float64)
    flags = np.zeros(n_points, dtype=np.uint8)
    base = point_offsets[:-1]
    points[base, 0], points[base, 1] = ex[:, 0], ey[:, 0]
    for seg in range(kmax):
        live = k > seg
        b = base[live] + 1 + 3 * seg
        nx = nxt[live, seg]
        li = np.nonzero(live)[0]
        points[b, 0], points[b, 1] = c0x[live, seg], c0y[live, seg]
        points[b + 1, 0], points[b + 1, 1] = c1x[live, seg], c1y[live, seg]
        points[b + 2, 0], points[b + 2, 1] = ex[li, nx], ey[li, nx]
        flags[b], flags[b + 1] = 1, 2

    pal = splitmix64(seed, n_paths * DRAWS_PER_PATH, PALETTE_SIZE * 2)
    rgb = pal[:PALETTE_SIZE]
    au = _unit(pal[PALETTE_SIZE:])
    palette = np.zeros((PALETTE_SIZE, 4), dtype=np.uint8)
    palette[:, 0] = (rgb & np.uint64(0xFF)).astype(np.uint8)
    palette[:, 1] = ((rgb >> np.uint64(8)) & np.uint64(0xFF)).astype(np.uint8)
    palette[:, 2] = ((rgb >> np.uint64(16)) & np.uint64(0xFF)).astype(np.uint8)
    opaque = ((rgb >> np.uint64(24)) & np.uint64(1)).astype(bool)
    palette[:, 3] = np.where(opaque, 255, 64 + np.minimum((au * 191).astype(np.int64), 190)).astype(np.uint8)

    fill_rules = np.where(np.arange(n_paths) % 2 == 1, FILL_RULE_EVEN_ODD, FILL_RULE_WINDING).astype(np.uint8)
    return FlatScene(points.astype(np.float32), flags, point_offsets.astype(np.uint32),
                     np.arange(n_paths + 1, dtype=np.uint32), fill_rules, pal_idx.astype(np.uint16), palette,
                     (0.0, 0.0, float(size), float(size)),
                     name or f"random{n_paths}@{size}", {"seed": seed, "size": size})


def tiger(size: int, even_odd_odd_paths: bool = False) -> tuple[FlatScene, tuple]:
    """The Ghostscript tiger framed like the reference demo: view box (0, 0, size, size) and the
    Camera::new_2d transform (demo/common/src/camera.rs:55-60,186-188; lib.rs:910-914). Returns the
    scene (in SVG user units) and the 2-D transform (m11, m12, m21, m22, tx, ty) in float32."""
    flat = FlatScene.load(TIGER_FIXTURE)
    svg_view_box = flat.view_box
    f = np.float32
    vw, vh = f(svg_view_box[2] - svg_view_box[0]), f(svg_view_box[3] - svg_view_box[1])
    scale = f(min(size, size)) * (f(1.0) / min(vw, vh))
    ox = f(size) * f(0.5) - vw * (scale * f(0.5))
    oy = f(size) * f(0.5) - vh * (scale * f(0.5))
    scene = flat.with_view_box((0.0, 0.0, float(size), float(size)))
    if even_odd_odd_paths:
        rules = scene.fill_rules.copy()
        rules[1::2] = FILL_RULE_EVEN_ODD
        scene = scene.with_fill_rules(rules)
    scene.name = f"tiger@{size}" + ("-evenodd" if even_odd_odd_paths else "")
    return scene, (float(scale), 0.0, 0.0, float(scale), float(ox), float(oy))


_GLYPHS = None


def _glyph_fixture():
    global _GLYPHS
    if _GLYPHS is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "roboto_glyphs.npz")
        _GLYPHS = {k: v for k, v in np.load(path).items()}
    return _GLYPHS


def text_page(n_glyphs: int = 10000, size: int = 2048, seed: int = 0x5EED0003, layout: str = "grid") -> FlatScene:
    """BASELINE.json configs[2] / SURVEY.md §8d config 3, outlines only: `n_glyphs` glyphs of Roboto Regular
    (tests/golden/roboto_glyphs.npz, built by tools/make_glyph_fixture.py), sizes uniform in {12..16} px, one path per
    glyph, black, winding rule, on a size x size page. layout "grid" (config 3): glyph ids uniform over the printable
    ASCII set, one glyph per cell of a ceil(sqrt(n))^2 grid; "lines": lines of random lower-case words, which puts
    neighbouring glyphs in the same tiles like running text does. Many tiny paths of quadratics: the opposite regime
    of the tiger. (Subpixel AA, stem darkening and the text filter of the reference's demo are §8 f3, not applied.)"""
    g = _glyph_fixture()
    upem = float(g["units_per_em"])
    codes, advances = g["codes"], g["advances"].astype(np.float64)
    letters = np.nonzero((codes >= ord("a")) & (codes <= ord("z")))[0]
    n_codes = len(codes)
    stream = splitmix64(seed, 0, 4 * n_glyphs + 4096)  # more draws than either layout can use
    cursor = [0]

    def rnd() -> int:
        cursor[0] += 1
        return int(stream[cursor[0] - 1])

    points, flags, contour_offsets, path_contour_offsets = [], [], [0], [0]

    def place(gi: int, x: float, y: float, scale: float):
        c0, c1 = int(g["glyph_contours"][gi]), int(g["glyph_contours"][gi + 1])
        for c in range(c0, c1):
            p0, p1 = int(g["contour_offsets"][c]), int(g["contour_offsets"][c + 1])
            pts = g["points"][p0:p1].astype(np.float64)
            out = np.empty_like(pts)
            out[:, 0] = x + pts[:, 0] * scale
            out[:, 1] = y - pts[:, 1] * scale
            # (a contour whose last segment is a curve ends on a copy of its first point, as pathfinder's
            # contours do: the implicit closing line is then empty)
            points.append(out)
            flags.append(g["point_flags"][p0:p1])
            contour_offsets.append(contour_offsets[-1] + len(out))
        path_contour_offsets.append(len(contour_offsets) - 1)

    if layout == "grid":
        side = int(np.ceil(np.sqrt(n_glyphs)))
        cell = size / side
        for i in range(n_glyphs):
            px = 12 + rnd() % 5
            gi = rnd() % n_codes
            place(gi, (i % side) * cell + 0.15 * cell, (i // side) * cell + 0.75 * cell, px / upem)
    elif layout == "lines":
        margin = 0.04 * size
        y = margin
        placed = 0
        while placed < n_glyphs:
            px = 12 + rnd() % 5
            scale = px / upem
            y += 1.3 * px
            if y > size - margin:  # page full: start over at the top, shifted, so any glyph count fits
                y = margin + 1.3 * px + 0.37 * px
            x = margin
            while placed < n_glyphs:
                word = 1 + rnd() % 9
                glyph_ids = [int(letters[rnd() % len(letters)]) if rnd() % 8 else rnd() % n_codes for _ in range(word)]
                word_width = float(sum(advances[i] for i in glyph_ids)) * scale
                if x + word_width > size - margin:
                    break
                for gi in glyph_ids:
                    if placed >= n_glyphs:
                        break
                    place(gi, x, y, scale)
                    placed += 1
                    x += float(advances[gi]) * scale
                x += float(g["space_advance"]) * scale
    else:
        raise ValueError(f"unknown layout {layout!r}")
    n_paths = len(path_contour_offsets) - 1
    return FlatScene(np.concatenate(points).astype(np.float32), np.concatenate(flags), contour_offsets, path_contour_offsets,
                     np.zeros(n_paths, np.uint8), np.zeros(n_paths, np.uint16), np.asarray([[0, 0, 0, 255]], np.uint8),
                     (0.0, 0.0, float(size), float(size)), f"text{n_glyphs}@{size}/{layout}", {"seed": seed, "size": size})


def text_page_subpixel(n_glyphs: int, size: int, layout: str = "grid"):
    """BASELINE.json configs[2] the way the reference's demo renders text with subpixel AA
    (demo/common/src/lib.rs:804-834): the page's glyphs, white, x scaled by 3, for a render target three times as
    wide as the page; one page-sized rectangle then samples that target through PatternFilter::Text. Returns the
    glyph scene (view box (0, 0, 3 * size, size)); `subpixel_scene` below assembles the whole Scene."""
    flat = text_page(n_glyphs, size, layout=layout)
    wide = flat.with_view_box((0.0, 0.0, 3.0 * size, float(size)))
    pts = wide.points.copy()
    pts[:, 0] = pts[:, 0] * np.float32(3.0)  # what Transform2F::from_scale(vec2f(3.0, 1.0)) does to a point (scene.rs:260-262)
    wide.points = pts
    wide.paint_colors = np.asarray([[255, 255, 255, 255]], np.uint8)  # the filter reads the red channel as coverage
    wide.paints = np.zeros_like(wide.paints)
    wide.name = f"text{n_glyphs}@{size}/subpixel"
    return wide


def subpixel_scene(wide: FlatScene, size: int, fg=(0.0, 0.0, 0.0), bg=(1.0, 1.0, 1.0),
                   kernel=(0.033165660, 0.102074051, 0.221434336, 0.286651906), gamma: bool = True):
    """The demo's wrapping (build_svg_tree, demo/common/src/lib.rs:804-834) as an api.Scene: render target 3 * size
    wide <- glyphs; page rectangle painted with the target as a pattern under PatternFilter::Text
    (defringing kernel DEFRINGING_KERNEL_CORE_GRAPHICS by default). The pattern transform scale(1/3, 1) makes page
    pixel x sample target texel 3x + 1 (its centre subpixel)."""
    from . import api
    scene = api.Scene()
    scene.set_view_box((0.0, 0.0, 3.0 * size, float(size)))
    target = scene.push_render_target(3 * size, size)
    scene.push_flat(wide)
    scene.pop_render_target()
    paint = scene.push_render_target_pattern(target, transform=(1.0 / 3.0, 0.0, 0.0, 1.0, 0.0, 0.0),
                                             text_filter={"fg": fg, "bg": bg, "kernel": kernel, "gamma": gamma})
    rect = np.asarray([[0, 0], [size, 0], [size, size], [0, size]], np.float32)
    scene.push_draw_path(rect, np.zeros(4, np.uint8), np.asarray([0, 4], np.uint32), paint)
    return scene
