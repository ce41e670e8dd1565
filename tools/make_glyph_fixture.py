#!/usr/bin/env python3
"""Builds tests/golden/roboto_glyphs.npz: the outlines (font units, TrueType quadratics) and advance widths of the
printable ASCII glyphs of the reference's resources/fonts/Roboto-Regular.ttf (Apache-2.0), read straight from the
`glyf` / `loca` / `cmap` / `hmtx` tables. pathfinder_b200.scenes.text_page() lays them out into the text-page scene of
BASELINE.json configs[2] (the outlines only: subpixel AA, stem darkening and the gamma LUT are SURVEY.md §8 f3).

The reference gets glyph outlines from font-kit 0.6.0 (text/src/lib.rs:80-160), which is not vendored: inputs derived
from fonts are parity-unpinned; the Scene built from this fixture is the comparison origin.

Usage: python tools/make_glyph_fixture.py [--check]
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FONT = "/root/reference/resources/fonts/Roboto-Regular.ttf"
OUT = os.path.join(ROOT, "tests", "golden", "roboto_glyphs.npz")


class Font:
    def __init__(self, data: bytes):
        self.d = data
        n = struct.unpack(">H", data[4:6])[0]
        self.tables = {}
        for i in range(n):
            tag, _chk, off, length = struct.unpack(">4sIII", data[12 + 16 * i:28 + 16 * i])
            self.tables[tag.decode()] = (off, length)
        head = self.tables["head"][0]
        self.units_per_em = struct.unpack(">H", data[head + 18:head + 20])[0]
        self.loca_long = struct.unpack(">h", data[head + 50:head + 52])[0] == 1
        self.num_glyphs = struct.unpack(">H", data[self.tables["maxp"][0] + 4:self.tables["maxp"][0] + 6])[0]
        self.n_hmetrics = struct.unpack(">H", data[self.tables["hhea"][0] + 34:self.tables["hhea"][0] + 36])[0]
        self.cmap = self._read_cmap()

    def _read_cmap(self):
        off = self.tables["cmap"][0]
        n = struct.unpack(">H", self.d[off + 2:off + 4])[0]
        best = None
        for i in range(n):
            platform, encoding, sub = struct.unpack(">HHI", self.d[off + 4 + 8 * i:off + 12 + 8 * i])
            fmt = struct.unpack(">H", self.d[off + sub:off + sub + 2])[0]
            if fmt == 4 and (platform, encoding) in ((3, 1), (0, 3), (0, 4), (0, 1)):
                best = off + sub
        assert best is not None, "no format-4 cmap subtable"
        o = best
        segx2 = struct.unpack(">H", self.d[o + 6:o + 8])[0]
        seg = segx2 // 2
        ends = struct.unpack(">%dH" % seg, self.d[o + 14:o + 14 + segx2])
        starts = struct.unpack(">%dH" % seg, self.d[o + 16 + segx2:o + 16 + 2 * segx2])
        deltas = struct.unpack(">%dh" % seg, self.d[o + 16 + 2 * segx2:o + 16 + 3 * segx2])
        ro_off = o + 16 + 3 * segx2
        range_offsets = struct.unpack(">%dH" % seg, self.d[ro_off:ro_off + segx2])
        mapping = {}
        for code in range(32, 127):
            for k in range(seg):
                if starts[k] <= code <= ends[k]:
                    if range_offsets[k] == 0:
                        gid = (code + deltas[k]) & 0xFFFF
                    else:
                        addr = ro_off + 2 * k + range_offsets[k] + 2 * (code - starts[k])
                        gid = struct.unpack(">H", self.d[addr:addr + 2])[0]
                        if gid:
                            gid = (gid + deltas[k]) & 0xFFFF
                    mapping[code] = gid
                    break
        return mapping

    def advance(self, gid):
        off = self.tables["hmtx"][0]
        k = min(gid, self.n_hmetrics - 1)
        return struct.unpack(">H", self.d[off + 4 * k:off + 4 * k + 2])[0]

    def glyph_range(self, gid):
        off = self.tables["loca"][0]
        if self.loca_long:
            a, b = struct.unpack(">II", self.d[off + 4 * gid:off + 4 * gid + 8])
        else:
            a, b = (2 * v for v in struct.unpack(">HH", self.d[off + 2 * gid:off + 2 * gid + 4]))
        g = self.tables["glyf"][0]
        return g + a, g + b

    def contours(self, gid, depth=0):
        """List of contours; a contour is a list of (x, y, on_curve) in font units."""
        a, b = self.glyph_range(gid)
        if a == b or depth > 4:
            return []
        d = self.d
        n_contours = struct.unpack(">h", d[a:a + 2])[0]
        p = a + 10
        if n_contours >= 0:
            ends = struct.unpack(">%dH" % n_contours, d[p:p + 2 * n_contours])
            p += 2 * n_contours
            n_instr = struct.unpack(">H", d[p:p + 2])[0]
            p += 2 + n_instr
            n_points = ends[-1] + 1 if n_contours else 0
            flags = []
            while len(flags) < n_points:
                f = d[p]; p += 1
                flags.append(f)
                if f & 8:
                    r = d[p]; p += 1
                    flags += [f] * r
            xs, x = [], 0
            for f in flags:
                if f & 2:
                    dx = d[p]; p += 1
                    x += dx if f & 16 else -dx
                elif not f & 16:
                    x += struct.unpack(">h", d[p:p + 2])[0]; p += 2
                xs.append(x)
            ys, y = [], 0
            for f in flags:
                if f & 4:
                    dy = d[p]; p += 1
                    y += dy if f & 32 else -dy
                elif not f & 32:
                    y += struct.unpack(">h", d[p:p + 2])[0]; p += 2
                ys.append(y)
            out, start = [], 0
            for e in ends:
                out.append([(xs[i], ys[i], bool(flags[i] & 1)) for i in range(start, e + 1)])
                start = e + 1
            return out
        out = []
        while True:  # composite glyph
            cflags, cgid = struct.unpack(">HH", d[p:p + 4]); p += 4
            if cflags & 1:
                dx, dy = struct.unpack(">hh", d[p:p + 4]); p += 4
            else:
                dx, dy = struct.unpack(">bb", d[p:p + 2]); p += 2
            m = [1.0, 0.0, 0.0, 1.0]
            if cflags & 8:
                s = struct.unpack(">h", d[p:p + 2])[0] / 16384.0; p += 2
                m = [s, 0.0, 0.0, s]
            elif cflags & 0x40:
                sx, sy = (v / 16384.0 for v in struct.unpack(">hh", d[p:p + 4])); p += 4
                m = [sx, 0.0, 0.0, sy]
            elif cflags & 0x80:
                m = [v / 16384.0 for v in struct.unpack(">hhhh", d[p:p + 8])]; p += 8
            assert cflags & 2, "composite glyphs positioned by point matching are not handled"
            for c in self.contours(cgid, depth + 1):
                out.append([(m[0] * x + m[2] * y + dx, m[1] * x + m[3] * y + dy, on) for x, y, on in c])
            if not cflags & 0x20:
                break
        return out


def to_quadratics(contour):
    """TrueType contour -> (points, flags) in pathfinder's layout: on-curve points with flag 0, quadratic control points
    with flag 1; implied on-curve points between consecutive off-curve points are made explicit."""
    n = len(contour)
    if n == 0:
        return [], []
    # FreeType's outline decomposition (font-kit's loader on Linux): start at the first point if it is on the curve,
    # else at the last point if that one is, else halfway between the two (csrc/font.cpp follows the same rule).
    if contour[0][2]:
        first, seq = contour[0][:2], contour[1:]
    elif contour[-1][2]:
        first, seq = contour[-1][:2], contour[:-1]
    else:
        (x0, y0, _), (x1, y1, _) = contour[0], contour[-1]
        first, seq = ((x0 + x1) / 2.0, (y0 + y1) / 2.0), contour
    pts, flags = [first], [0]
    pending = None
    for x, y, on in seq:
        if on:
            if pending is not None:
                pts.append(pending); flags.append(1)
                pending = None
            pts.append((x, y)); flags.append(0)
        else:
            if pending is not None:
                mid = ((pending[0] + x) / 2.0, (pending[1] + y) / 2.0)
                pts.append(pending); flags.append(1)
                pts.append(mid); flags.append(0)
            pending = (x, y)
    if pending is not None:  # the closing segment is a curve back to the first point
        pts.append(pending); flags.append(1)
        pts.append(first); flags.append(0)
    return pts, flags


def build():
    font = Font(open(FONT, "rb").read())
    codes, advances, points, flags, contour_offsets, glyph_contours = [], [], [], [], [0], [0]
    for code in range(33, 127):
        gid = font.cmap.get(code, 0)
        assert gid, f"no glyph for {chr(code)!r}"
        codes.append(code)
        advances.append(font.advance(gid))
        for c in font.contours(gid):
            p, f = to_quadratics(c)
            if len(p) < 3:
                continue
            points += p
            flags += f
            contour_offsets.append(len(points))
        glyph_contours.append(len(contour_offsets) - 1)
    return dict(codes=np.asarray(codes, np.uint16), advances=np.asarray(advances, np.uint16),
                space_advance=np.asarray(font.advance(font.cmap[32]), np.uint16),
                units_per_em=np.asarray(font.units_per_em, np.uint16),
                points=np.asarray(points, np.float32).reshape(-1, 2), point_flags=np.asarray(flags, np.uint8),
                contour_offsets=np.asarray(contour_offsets, np.uint32), glyph_contours=np.asarray(glyph_contours, np.uint32))


def main():
    data = build()
    if "--check" in sys.argv[1:]:
        have = np.load(OUT)
        same = all(np.array_equal(have[k], v) for k, v in data.items())
        print("roboto_glyphs.npz:", "identical" if same else "DIFFERENT")
        sys.exit(0 if same else 1)
    np.savez_compressed(OUT, **data)
    print(f"{len(data['codes'])} glyphs, {len(data['contour_offsets']) - 1} contours, {len(data['points'])} points, "
          f"unitsPerEm {int(data['units_per_em'])} -> {OUT} ({os.path.getsize(OUT)} bytes)")


if __name__ == "__main__":
    main()
