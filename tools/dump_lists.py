#!/usr/bin/env python3
"""Dumps a scene and its tile / fill / z-buffer lists as text, for comparison with the reference.

  python tools/dump_lists.py tiger 1024 --source oracle --out tiger1024.lists --scene-out tiger1024.scene
  python tools/dump_lists.py random:3000:7 1024 --source cuda --out random.lists      (needs a GPU)

`--scene-out` writes the exact input (post-loader outlines, the comparison origin of SURVEY.md §8c) in the
format tools/reference_dump/main.rs reads; running that program inside a checkout of the reference prints
the same list format from `Scene::build_into_vector` at the D3D9 level with `SequentialExecutor`.
tools/diff_lists.py compares two list files after canonicalising alpha-tile ids.

List format, one record per line:
  fill <from_x> <from_y> <to_x> <to_y> <alpha_tile_id>      (LineSegmentU16 in 8.8 fixed point, gpu_data.rs Fill)
  tile <tile_x> <tile_y> <alpha_tile_id> <path_id> <color> <ctrl> <backdrop>   (TileObjectPrimitive)
  z <width> <height> <v0> <v1> ...                          (DrawTilesD3D9 z-buffer texels, row-major)
  clip <dest_tile_id> <dest_backdrop> <src_tile_id> <src_backdrop>             (Clip, DrawTileBatchD3D9.clips)
Scene format:
  viewbox <min_x> <min_y> <max_x> <max_y>
  transform <m11> <m12> <m21> <m22> <tx> <ty>               (BuildOptions transform, Transform2F row-major)
  dilation <x> <y>                                           (BuildOptions dilation; absent = zero)
  paint <r> <g> <b> <a>
  clippath <fill_rule> <contour count>                       (clip paths first, ids in file order)
  path <fill_rule: 0 winding | 1 even-odd> <paint index> <contour count> [<clip path id> | -1]
  contour <point count>
  p <x> <y> <flags>                                          (f32 as hex bits; flags: 1 = control 0, 2 = control 1)
"""
import argparse
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathfinder_b200 import scenes  # noqa: E402


def load(spec, size):
    if spec == "tiger":
        return scenes.tiger(size)
    if spec == "clips":
        from tests.test_parity_gpu import clip_scene
        return clip_scene(size), None
    if spec.startswith("random:"):
        _, n, seed = spec.split(":")
        return scenes.random_paths(int(n), size, int(seed)), None
    if spec.startswith("text:"):
        _, n, layout = spec.split(":")
        return scenes.text_page(int(n), size, layout=layout), None
    raise SystemExit(f"unknown scene {spec!r} (tiger | clips | random:<paths>:<seed> | text:<glyphs>:<grid|lines>)")


def bits(v):
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(v)))[0]


def write_scene(flat, xf, path, dilation=(0.0, 0.0)):
    with open(path, "w") as f:
        f.write("viewbox " + " ".join(bits(v) for v in flat.view_box) + "\n")
        t = (1, 0, 0, 1, 0, 0) if xf is None else xf
        f.write("transform " + " ".join(bits(v) for v in t) + "\n")
        if dilation[0] != 0.0 or dilation[1] != 0.0:
            f.write("dilation " + " ".join(bits(v) for v in dilation) + "\n")
        for c in np.asarray(flat.paint_colors).reshape(-1, 4):
            f.write("paint %d %d %d %d\n" % tuple(int(v) for v in c))
        def contours(c0, c1):
            for c in range(int(c0), int(c1)):
                p0, p1 = int(flat.contour_offsets[c]), int(flat.contour_offsets[c + 1])
                f.write("contour %d\n" % (p1 - p0))
                for k in range(p0, p1):
                    f.write("p %s %s %d\n" % (bits(flat.points[k][0]), bits(flat.points[k][1]), int(flat.point_flags[k])))

        for (c0, c1), rule in zip(flat.clip_contour_ranges, flat.clip_fill_rules):
            f.write("clippath %d %d\n" % (int(rule), int(c1 - c0)))
            contours(c0, c1)
        paints = flat.paints  # table indices: the reader pushes the table in order
        for i, (c0, c1) in enumerate(flat.contour_ranges()):
            clip = int(flat.draw_clip_paths[i])
            f.write("path %d %d %d %d\n" % (int(flat.fill_rules[i]), int(paints[i]), int(c1 - c0), -1 if clip == 0xFFFFFFFF else clip))
            contours(c0, c1)


def write_lists(fills, tiles, z, path, clips=()):
    with open(path, "w") as f:
        for r in fills:
            f.write("fill %d %d %d %d %d\n" % (r["from_x"], r["from_y"], r["to_x"], r["to_y"], r["link"]))
        for r in tiles:
            f.write("tile %d %d %d %d %d %d %d\n" % (r["tile_x"], r["tile_y"], r["alpha_tile_id"], r["path_id"],
                                                   r["color"], r["ctrl"], r["backdrop"]))
        z = np.asarray(z)
        f.write("z %d %d %s\n" % (z.shape[1], z.shape[0], " ".join(str(int(v)) for v in z.reshape(-1))))
        for r in clips:
            f.write("clip %d %d %d %d\n" % (r["dest_tile_id"], r["dest_backdrop"], r["src_tile_id"], r["src_backdrop"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("size", type=int)
    ap.add_argument("--source", default="oracle", choices=["oracle", "cuda"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--scene-out")
    ap.add_argument("--dilation", type=float, nargs=2, default=(0.0, 0.0), metavar=("X", "Y"),
                    help="BuildOptions dilation in pixels, applied after the transform (stem darkening)")
    args = ap.parse_args()
    flat, xf = load(args.scene, args.size)
    dilation = (float(np.float32(args.dilation[0])), float(np.float32(args.dilation[1])))
    if args.scene_out:
        write_scene(flat, xf, args.scene_out, dilation)
    if args.source == "oracle":
        from tests import helpers as H
        b = H.oracle_build(flat, xf, dilation=dilation)
        write_lists(b.fills, b.tiles, b.z_buffer, args.out, b.clips)
    else:
        from tests import helpers as H
        r, _ = H.cuda_render(flat, xf, size=(args.size, args.size), debug=True, dilation=dilation)
        z, _rect = r.debug_z_buffer()
        write_lists(r.debug_fills(), r.debug_tiles(), z, args.out, r.debug_clips())


if __name__ == "__main__":
    main()
