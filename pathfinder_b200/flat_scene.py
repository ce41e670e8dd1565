"""Flat (numpy) scene container: the comparison origin for parity (SURVEY.md §8c — the Scene,
i.e. post-loader outlines, is where the CUDA path and the oracle start from).

Paths own contiguous contour ranges; contours own contiguous point ranges. Point flags follow
`content/src/outline.rs` PointFlags (0 on-curve, 1 CONTROL_POINT_0, 2 CONTROL_POINT_1)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FILL_RULE_WINDING = 0
FILL_RULE_EVEN_ODD = 1
NO_CLIP = 0xFFFFFFFF


@dataclass
class FlatScene:
    points: np.ndarray            # (n_points, 2) float32
    point_flags: np.ndarray       # (n_points,) uint8
    contour_offsets: np.ndarray   # (n_contours + 1,) uint32 -> point index
    path_contour_offsets: np.ndarray  # (n_paths + 1,) uint32 -> contour index (draw paths)
    fill_rules: np.ndarray        # (n_paths,) uint8
    paints: np.ndarray            # (n_paths,) uint16 paint id
    paint_colors: np.ndarray      # (n_paints, 4) uint8 RGBA
    view_box: tuple               # (min_x, min_y, max_x, max_y)
    name: str = ""
    meta: dict = field(default_factory=dict)
    # Clip paths (optional): their contours follow the draw paths' contours in the same pools.
    clip_contour_ranges: np.ndarray | None = None  # (n_clip_paths, 2) uint32 contour ranges
    clip_fill_rules: np.ndarray | None = None      # (n_clip_paths,) uint8
    draw_clip_paths: np.ndarray | None = None      # (n_paths,) uint32 clip path id, NO_CLIP = none

    def __post_init__(self):
        self.points = np.ascontiguousarray(self.points, dtype=np.float32).reshape(-1, 2)
        self.point_flags = np.ascontiguousarray(self.point_flags, dtype=np.uint8)
        self.contour_offsets = np.ascontiguousarray(self.contour_offsets, dtype=np.uint32)
        self.path_contour_offsets = np.ascontiguousarray(self.path_contour_offsets, dtype=np.uint32)
        self.fill_rules = np.ascontiguousarray(self.fill_rules, dtype=np.uint8)
        self.paints = np.ascontiguousarray(self.paints, dtype=np.uint16)
        self.paint_colors = np.ascontiguousarray(self.paint_colors, dtype=np.uint8).reshape(-1, 4)
        self.view_box = tuple(float(v) for v in self.view_box)
        assert len(self.points) == len(self.point_flags)
        assert int(self.contour_offsets[-1]) == len(self.points)
        assert len(self.fill_rules) == self.n_paths and len(self.paints) == self.n_paths
        if self.clip_contour_ranges is None:
            assert int(self.path_contour_offsets[-1]) == len(self.contour_offsets) - 1
            self.clip_contour_ranges = np.zeros((0, 2), dtype=np.uint32)
            self.clip_fill_rules = np.zeros(0, dtype=np.uint8)
        self.clip_contour_ranges = np.ascontiguousarray(self.clip_contour_ranges, dtype=np.uint32).reshape(-1, 2)
        self.clip_fill_rules = np.ascontiguousarray(self.clip_fill_rules, dtype=np.uint8)
        if self.draw_clip_paths is None:
            self.draw_clip_paths = np.full(self.n_paths, NO_CLIP, dtype=np.uint32)
        self.draw_clip_paths = np.ascontiguousarray(self.draw_clip_paths, dtype=np.uint32)
        assert len(self.draw_clip_paths) == self.n_paths and len(self.clip_fill_rules) == self.n_clip_paths

    @property
    def n_paths(self) -> int:
        return len(self.path_contour_offsets) - 1

    @property
    def n_clip_paths(self) -> int:
        return len(self.clip_contour_ranges)

    @property
    def n_contours(self) -> int:
        return len(self.contour_offsets) - 1

    def palette(self):
        """(paint id per path, colour table) as Scene::push_paint assigns them: equal colours share one id,
        ids in order of first appearance in the table (Palette::push_paint, renderer/src/paint.rs:115-131)."""
        colors = np.ascontiguousarray(self.paint_colors, dtype=np.uint8).reshape(-1, 4)
        keys = colors.view(np.uint32).reshape(-1)
        _, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first)                 # unique colours in order of first appearance
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        remap = rank[inverse].astype(np.uint16)   # table index -> deduplicated paint id
        return remap[self.paints], colors[np.sort(first)]

    def contour_ranges(self) -> np.ndarray:
        return np.stack([self.path_contour_offsets[:-1], self.path_contour_offsets[1:]], axis=1).astype(np.uint32)

    def with_view_box(self, view_box) -> "FlatScene":
        return FlatScene(self.points, self.point_flags, self.contour_offsets, self.path_contour_offsets,
                         self.fill_rules, self.paints, self.paint_colors, view_box, self.name, dict(self.meta))

    def with_fill_rules(self, fill_rules) -> "FlatScene":
        return FlatScene(self.points, self.point_flags, self.contour_offsets, self.path_contour_offsets,
                         fill_rules, self.paints, self.paint_colors, self.view_box, self.name, dict(self.meta))

    def save(self, path: str) -> None:
        np.savez_compressed(path, points=self.points, point_flags=self.point_flags,
                            contour_offsets=self.contour_offsets,
                            path_contour_offsets=self.path_contour_offsets, fill_rules=self.fill_rules,
                            paints=self.paints, paint_colors=self.paint_colors,
                            view_box=np.asarray(self.view_box, dtype=np.float32),
                            name=np.asarray(self.name))

    @staticmethod
    def load(path: str) -> "FlatScene":
        z = np.load(path)
        return FlatScene(z["points"], z["point_flags"], z["contour_offsets"], z["path_contour_offsets"],
                         z["fill_rules"], z["paints"], z["paint_colors"], tuple(z["view_box"].tolist()),
                         str(z["name"]))


class SceneBuilderPy:
    """Incremental builder of a FlatScene with a canvas-like path API (move_to / line_to / ...),
    mirroring how `content/src/outline.rs` Contour::push_endpoint / push_quadratic / push_cubic lay
    points out."""

    def __init__(self, view_box):
        self.view_box = tuple(view_box)
        self._points: list = []
        self._flags: list = []
        self._contour_offsets = [0]
        self._path_contour_offsets = [0]
        self._fill_rules: list = []
        self._paints: list = []
        self._paint_colors: list = []
        self._paint_cache: dict = {}
        self._open = False
        self._clip_paths: list = []   # (points, flags, local contour offsets, fill rule)
        self._draw_clip: list = []

    def paint(self, rgba) -> int:
        key = tuple(int(v) for v in rgba)
        if key not in self._paint_cache:
            self._paint_cache[key] = len(self._paint_colors)
            self._paint_colors.append(key)
        return self._paint_cache[key]

    def move_to(self, x, y):
        self._end_contour()
        self._points.append((x, y))
        self._flags.append(0)
        self._open = True

    def line_to(self, x, y):
        self._points.append((x, y))
        self._flags.append(0)

    def quad_to(self, cx, cy, x, y):
        self._points += [(cx, cy), (x, y)]
        self._flags += [1, 0]

    def cubic_to(self, c0x, c0y, c1x, c1y, x, y):
        self._points += [(c0x, c0y), (c1x, c1y), (x, y)]
        self._flags += [1, 2, 0]

    def close(self):
        self._end_contour()

    def add_outline(self, points, point_flags, contour_offsets):
        """Appends whole contours (points + PointFlags, contour i = points [offsets[i], offsets[i + 1])) to the
        path being built — the form PFSvgPathDataToOutline / PFOutlineStrokeToFill return."""
        self._end_contour()
        pts = np.asarray(points, dtype=np.float32).reshape(-1, 2)
        for i in range(len(contour_offsets) - 1):
            a, b = int(contour_offsets[i]), int(contour_offsets[i + 1])
            if b > a:
                self._points += [(float(x), float(y)) for x, y in pts[a:b]]
                self._flags += [int(f) for f in point_flags[a:b]]
                self._contour_offsets.append(len(self._points))

    def _end_contour(self):
        if self._open and len(self._points) > self._contour_offsets[-1]:
            self._contour_offsets.append(len(self._points))
        self._open = False

    def end_path(self, rgba, fill_rule=FILL_RULE_WINDING, clip=NO_CLIP):
        self._end_contour()
        self._path_contour_offsets.append(len(self._contour_offsets) - 1)
        self._fill_rules.append(fill_rule)
        self._paints.append(self.paint(rgba))
        self._draw_clip.append(clip)

    def end_clip_path(self, fill_rule=FILL_RULE_WINDING) -> int:
        """Turns the contours drawn since the last end_path / end_clip_path into a clip path
        (Scene::push_clip_path) and returns its id, to be passed as end_path(..., clip=id)."""
        self._end_contour()
        c0 = self._path_contour_offsets[-1]
        p0 = self._contour_offsets[c0]
        local = [o - p0 for o in self._contour_offsets[c0:]]
        self._clip_paths.append((self._points[p0:], self._flags[p0:], local, fill_rule))
        del self._points[p0:], self._flags[p0:], self._contour_offsets[c0 + 1:]
        return len(self._clip_paths) - 1

    def finish(self, name="") -> FlatScene:
        self._end_contour()
        points, flags, contour_offsets = list(self._points), list(self._flags), list(self._contour_offsets)
        clip_ranges, clip_rules = [], []
        for cp_points, cp_flags, local, rule in self._clip_paths:  # clip contours follow the draw contours
            base, c0 = len(points), len(contour_offsets) - 1
            points += cp_points
            flags += cp_flags
            contour_offsets += [base + o for o in local[1:]]
            clip_ranges.append((c0, len(contour_offsets) - 1))
            clip_rules.append(rule)
        return FlatScene(np.asarray(points, dtype=np.float32).reshape(-1, 2), flags,
                         contour_offsets, self._path_contour_offsets, self._fill_rules, self._paints,
                         np.asarray(self._paint_colors, dtype=np.uint8).reshape(-1, 4), self.view_box, name,
                         clip_contour_ranges=np.asarray(clip_ranges, dtype=np.uint32).reshape(-1, 2) if clip_ranges else None,
                         clip_fill_rules=np.asarray(clip_rules, dtype=np.uint8) if clip_ranges else None,
                         draw_clip_paths=np.asarray(self._draw_clip, dtype=np.uint32) if clip_ranges else None)
