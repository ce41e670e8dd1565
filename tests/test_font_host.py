"""The TrueType reader on the host: csrc/font.cpp behind PFFont*, checked against the glyph
fixture that the independent Python reader (tools/make_glyph_fixture.py) produced from the same font file, against
hostile inputs, and through FontContext::push_glyph's transform chain (text/src/lib.rs:118-146)."""
import os
import struct

import numpy as np
import pytest

from pathfinder_b200 import _lib as L
from pathfinder_b200 import api, scenes

FONT = "/root/reference/resources/fonts/Roboto-Regular.ttf"
needs_font = pytest.mark.skipif(not os.path.exists(FONT), reason="the reference checkout (and its fonts) is only present in the build container")


@pytest.fixture(scope="module")
def font():
    return api.Font.from_path(FONT)


@needs_font
def test_glyphs_match_the_fixture(font):
    g = scenes._glyph_fixture()
    assert font.units_per_em == int(g["units_per_em"]) == 2048
    assert font.advance(font.glyph_for_char(" ")) == float(g["space_advance"])
    for k, code in enumerate(g["codes"]):
        gid = font.glyph_for_char(chr(int(code)))
        assert gid != 0, chr(int(code))
        assert font.advance(gid) == float(g["advances"][k])
        pts, flags, offsets = font.outline(gid)
        c0, c1 = int(g["glyph_contours"][k]), int(g["glyph_contours"][k + 1])
        p0, p1 = int(g["contour_offsets"][c0]), int(g["contour_offsets"][c1])
        assert np.array_equal(offsets, g["contour_offsets"][c0:c1 + 1] - p0), chr(int(code))
        assert pts.tobytes() == g["points"][p0:p1].tobytes(), chr(int(code))
        assert np.array_equal(flags, g["point_flags"][p0:p1])


@needs_font
def test_lookups_outside_the_font(font):
    assert font.glyph_for_char(0x10FFFF) == 0 and font.glyph_for_char(0xE000) == 0
    assert font.advance(font.glyph_count + 5) == 0.0
    pts, flags, offsets = font.outline(font.glyph_for_char(" "))       # no contours: an empty outline, not an error
    assert len(pts) == 0 and list(offsets) == [0]
    with pytest.raises(L.PathfinderCudaError):
        font.outline(font.glyph_count)


@needs_font
def test_composite_glyphs_are_resolved(font):
    """e-acute is a composite of 'e' and the acute accent: its outline is the base glyph's contours followed by
    the accent's, moved by the component offset."""
    base = font.outline(font.glyph_for_char("e"))
    composite = font.outline(font.glyph_for_char("é"))
    n_base = len(base[0])
    assert len(composite[2]) > len(base[2]) and len(composite[0]) > n_base
    assert composite[0][:n_base].tobytes() == base[0].tobytes()
    accent = composite[0][n_base:]
    assert accent[:, 1].min() > base[0][:, 1].max() * 0.9               # the accent sits above the letter


@needs_font
def test_push_glyph_transform_chain(font):
    """An 'H' at 16 px placed at (100, 200): y flips (font units are y-up), the stems are vertical, and the cap
    height is Roboto's 1456 / 2048 em."""
    gid = font.glyph_for_char("H")
    pts, flags, offsets = font.glyph_outline_at(gid, (100.0, 200.0), 16.0)
    assert not flags.any() and len(offsets) == 2
    assert abs(pts[:, 1].max() - 200.0) < 1e-4                          # the baseline
    assert abs((200.0 - pts[:, 1].min()) - 16.0 * 1456 / 2048) < 1e-3   # cap height
    assert pts[:, 0].min() > 100.0 and pts[:, 0].max() < 100.0 + 16.0 * font.advance(gid) / 2048
    # a render transform composes on the left
    t = api.Transform2F(2.0, 0.0, 0.0, 2.0, 5.0, -3.0)
    scaled, _, _ = font.glyph_outline_at(gid, (100.0, 200.0), 16.0, t)
    assert np.allclose(scaled, pts * 2.0 + np.array([5.0, -3.0], np.float32), atol=1e-3)


def test_garbage_is_refused():
    with pytest.raises(L.PathfinderCudaError):
        api.Font(b"")
    with pytest.raises(L.PathfinderCudaError):
        api.Font(b"OTTO" + bytes(64))                                    # CFF-flavoured: not read
    with pytest.raises(L.PathfinderCudaError):
        api.Font(struct.pack(">IHHHH", 0x00010000, 200, 0, 0, 0))        # a table directory that is not there


@needs_font
def test_truncated_and_corrupted_fonts_never_fault():
    """Every read is bounds-checked: prefixes of the file and files with flipped bytes either fail to open or
    answer every query (possibly with an error) without touching memory outside the copy."""
    data = open(FONT, "rb").read()
    lib = L.lib()

    def poke(blob):
        h = lib.PFFontCreateFromBytes(blob, len(blob))
        if not h:
            return 0
        n = lib.PFFontGetGlyphCount(h)
        for gid in list(range(0, min(n, 40))) + [n - 1, n, 0xFFFF]:
            lib.PFFontGetGlyphAdvance(h, gid)
            o = lib.PFFontGetGlyphOutline(h, gid)
            if o:
                lib.PFOutlineDestroy(o)
        for code in (0, 65, 0xE9, 0xFFFF, 0x1F600):
            lib.PFFontGetGlyphForCodepoint(h, code)
        lib.PFFontDestroy(h)
        return 1

    opened = sum(poke(data[:n]) for n in list(range(0, 600, 7)) + list(range(600, len(data), 9973)))
    assert opened == 0 or opened < 60          # a prefix that cuts a table off does not open
    rng = np.random.default_rng(5)
    for _ in range(60):
        blob = bytearray(data)
        for pos in rng.integers(0, len(blob), 200):
            blob[pos] ^= int(rng.integers(1, 256))
        poke(bytes(blob))
    # damage aimed at the structures the reader trusts most: table directory, loca, glyph headers
    for lo, hi in ((0, 300), (300, 4000)):
        for _ in range(40):
            blob = bytearray(data)
            for pos in rng.integers(lo, hi, 12):
                blob[pos] ^= int(rng.integers(1, 256))
            poke(bytes(blob))
    assert poke(data) == 1


@needs_font
def test_push_text_builds_the_same_scene_as_the_fixture_layout(font, area_lut):
    """Scene.push_text (advance-only layout) -> host SceneBuilder: the batch the renderer would receive has one path
    per inked glyph, and the oracle renders the same outlines to the ink their contours enclose."""
    from pathfinder_b200.flat_scene import FlatScene
    from tests import helpers as H
    from tests.test_scene_host import collect
    text = "Hamburgefonstiv 0123"
    scene = api.Scene()
    scene.set_view_box((0, 0, 256, 64))
    paint = scene.push_paint((0, 0, 0, 255))
    end_x = scene.push_text(font, text, (8.0, 40.0), 20.0, paint)
    inked = [c for c in text if c != " "]
    assert scene.draw_path_count() == len(inked)
    want = 8.0 + sum(font.advance(font.glyph_for_char(c)) for c in text) * 20.0 / 2048
    assert abs(end_x - want) < 1e-2
    cmds = collect(scene, api.BuildOptions())
    draw = [c for c in cmds if c["kind"] == "DrawTilesD3D11"][0]
    assert draw["path_count"] == len(inked)

    # the same outlines as a FlatScene through the oracle
    pts, flags, offs, path_offs = [], [], [0], [0]
    x = np.float32(8.0)
    for ch in text:
        g = font.glyph_for_char(ch)
        p, f, o = font.glyph_outline_at(g, (x, np.float32(40.0)), 20.0)
        x = np.float32(x + np.float32(font.advance(g)) * (np.float32(20.0) / np.float32(2048)))
        if len(p) == 0:
            continue
        base = offs[-1]
        pts.append(p); flags.append(f); offs += [int(v) + base for v in o[1:]]
        path_offs.append(len(offs) - 1)
    n = len(path_offs) - 1
    flat = FlatScene(np.concatenate(pts), np.concatenate(flags), offs, path_offs, np.zeros(n, np.uint8),
                     np.zeros(n, np.uint16), np.asarray([[0, 0, 0, 255]], np.uint8), (0.0, 0.0, 256.0, 64.0), "text")
    built = H.oracle_build(flat, None)
    assert draw["tile_count"] == built.bbox_tile_count and draw["segment_count"] == built.input_segment_count
    img = built.render(area_lut, 256, 64)
    assert 150 < img[:, :, 3].astype(np.float64).sum() / 255.0 < 1500   # some ink, not a filled page
