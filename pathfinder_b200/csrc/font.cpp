// pathfinder_b200/csrc/font.cpp — a small TrueType reader for the text path.
//
// The reference gets glyph ids, advances and outlines from font-kit 0.6.0 (text/src/lib.rs:80-160:
// `font.outline(glyph_id, hinting_options, &mut OutlinePathBuilder)` with HintingOptions::None), a dependency that
// is not vendored in the reference checkout; PARITY UNPINNED upstream of the Scene. This file restates the
// published format instead (OpenType 1.8 `cmap` formats 4 and 12, `head`, `maxp`, `hhea`, `hmtx`, `loca`, `glyf`
// with simple and composite glyphs) and turns TrueType contours into pathfinder's point lists the way FreeType's
// outline decomposition — font-kit's loader on Linux — walks them: start at the first point if it is on the curve,
// else at the last point if that one is, else at the midpoint of the two; consecutive off-curve points imply an
// on-curve point halfway. Every read is bounds-checked: a truncated or corrupt file yields an error, never a fault.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pf_cuda.h"
#include "outline.h"

namespace pf {
void set_last_error(const std::string &msg);
}

struct PFFont {
    std::vector<uint8_t> data;
    uint32_t glyf = 0, glyf_len = 0, loca = 0, loca_len = 0, hmtx = 0, hmtx_len = 0;
    uint32_t cmap4 = 0, cmap12 = 0; // offsets of the chosen subtables (0 = absent)
    uint32_t units_per_em = 0, glyph_count = 0, h_metrics = 0;
    bool loca_long = false;
};

namespace {

struct Bad {}; // thrown by the checked readers

struct Reader {
    const std::vector<uint8_t> &d;
    void need(size_t off, size_t n) const {
        if (off > d.size() || n > d.size() - off) throw Bad{};
    }
    uint8_t u8(size_t off) const {
        need(off, 1);
        return d[off];
    }
    uint16_t u16(size_t off) const {
        need(off, 2);
        return (uint16_t)(d[off] << 8 | d[off + 1]);
    }
    int16_t i16(size_t off) const { return (int16_t)u16(off); }
    uint32_t u32(size_t off) const {
        need(off, 4);
        return (uint32_t)d[off] << 24 | (uint32_t)d[off + 1] << 16 | (uint32_t)d[off + 2] << 8 | d[off + 3];
    }
};

bool find_table(const Reader &r, const char *tag, uint32_t &off, uint32_t &len) {
    const uint16_t n = r.u16(4);
    for (uint16_t i = 0; i < n; i++) {
        const size_t rec = 12 + 16 * (size_t)i;
        r.need(rec, 16);
        if (memcmp(&r.d[rec], tag, 4) == 0) {
            off = r.u32(rec + 8);
            len = r.u32(rec + 12);
            r.need(off, len);
            return true;
        }
    }
    return false;
}

struct GlyphPoint {
    float x, y;
    bool on_curve;
};
using GlyphContour = std::vector<GlyphPoint>;

constexpr int MAX_COMPONENT_DEPTH = 8;
constexpr size_t MAX_GLYPH_POINTS = 1u << 20; // composite glyphs of a hostile file cannot blow up memory
constexpr size_t MAX_COMPONENT_VISITS = 4096;  // ... nor recurse without end through empty components

void glyph_range(const PFFont &f, const Reader &r, uint32_t gid, uint32_t &begin, uint32_t &end) {
    if (gid >= f.glyph_count) throw Bad{};
    if (f.loca_long) {
        if ((uint64_t)4 * (gid + 2) > f.loca_len) throw Bad{};
        begin = r.u32(f.loca + 4 * (size_t)gid);
        end = r.u32(f.loca + 4 * (size_t)gid + 4);
    } else {
        if ((uint64_t)2 * (gid + 2) > f.loca_len) throw Bad{};
        begin = 2u * r.u16(f.loca + 2 * (size_t)gid);
        end = 2u * r.u16(f.loca + 2 * (size_t)gid + 2);
    }
    if (begin > end || end > f.glyf_len) throw Bad{};
}

void read_contours(const PFFont &f, const Reader &r, uint32_t gid, int depth, std::vector<GlyphContour> &out,
                   size_t &total_points) {
    uint32_t begin, end;
    glyph_range(f, r, gid, begin, end);
    if (begin == end) return; // no outline (space)
    if (depth > MAX_COMPONENT_DEPTH) throw Bad{};
    size_t p = (size_t)f.glyf + begin;
    const size_t limit = (size_t)f.glyf + end;
    auto need = [&](size_t n) {
        if (p > limit || n > limit - p) throw Bad{};
    };
    need(10);
    const int16_t n_contours = r.i16(p);
    p += 10;
    if (n_contours >= 0) {
        need(2 * (size_t)n_contours + 2);
        std::vector<uint16_t> ends((size_t)n_contours);
        for (int c = 0; c < n_contours; c++) {
            ends[c] = r.u16(p + 2 * (size_t)c);
            if (c > 0 && ends[c] < ends[c - 1]) throw Bad{};
        }
        p += 2 * (size_t)n_contours;
        const uint16_t n_instructions = r.u16(p);
        p += 2;
        need(n_instructions);
        p += n_instructions;
        const size_t n_points = n_contours ? (size_t)ends.back() + 1 : 0;
        total_points += n_points;
        if (total_points > MAX_GLYPH_POINTS) throw Bad{};
        std::vector<uint8_t> flags;
        flags.reserve(n_points);
        while (flags.size() < n_points) {
            need(1);
            const uint8_t fl = r.u8(p++);
            flags.push_back(fl);
            if (fl & 8) { // REPEAT_FLAG
                need(1);
                uint8_t repeat = r.u8(p++);
                while (repeat-- && flags.size() < n_points) flags.push_back(fl);
            }
        }
        std::vector<int32_t> xs(n_points), ys(n_points);
        int32_t v = 0;
        for (size_t i = 0; i < n_points; i++) {
            const uint8_t fl = flags[i];
            if (fl & 2) { // X_SHORT_VECTOR: one byte, sign in bit 4
                need(1);
                const int32_t d = r.u8(p++);
                v += (fl & 16) ? d : -d;
            } else if (!(fl & 16)) { // two bytes unless "same as previous"
                need(2);
                v += r.i16(p);
                p += 2;
            }
            xs[i] = v;
        }
        v = 0;
        for (size_t i = 0; i < n_points; i++) {
            const uint8_t fl = flags[i];
            if (fl & 4) {
                need(1);
                const int32_t d = r.u8(p++);
                v += (fl & 32) ? d : -d;
            } else if (!(fl & 32)) {
                need(2);
                v += r.i16(p);
                p += 2;
            }
            ys[i] = v;
        }
        size_t start = 0;
        for (int c = 0; c < n_contours; c++) {
            GlyphContour contour;
            for (size_t i = start; i <= ends[c]; i++)
                contour.push_back(GlyphPoint{(float)xs[i], (float)ys[i], (flags[i] & 1) != 0});
            start = (size_t)ends[c] + 1;
            out.push_back(std::move(contour));
        }
        return;
    }
    // Composite glyph: components placed by an offset and an optional 2x2 matrix (F2Dot14).
    for (;;) {
        need(4);
        const uint16_t cflags = r.u16(p);
        const uint16_t component = r.u16(p + 2);
        p += 4;
        float dx, dy;
        if (cflags & 1) { // ARG_1_AND_2_ARE_WORDS
            need(4);
            dx = (float)r.i16(p), dy = (float)r.i16(p + 2);
            p += 4;
        } else {
            need(2);
            dx = (float)(int8_t)r.u8(p), dy = (float)(int8_t)r.u8(p + 1);
            p += 2;
        }
        if (!(cflags & 2)) throw Bad{}; // ARGS_ARE_XY_VALUES unset: point matching is not handled
        float m[4] = {1.0f, 0.0f, 0.0f, 1.0f};
        const float k = 1.0f / 16384.0f;
        if (cflags & 8) { // WE_HAVE_A_SCALE
            need(2);
            m[0] = m[3] = (float)r.i16(p) * k;
            p += 2;
        } else if (cflags & 0x40) { // WE_HAVE_AN_X_AND_Y_SCALE
            need(4);
            m[0] = (float)r.i16(p) * k, m[3] = (float)r.i16(p + 2) * k;
            p += 4;
        } else if (cflags & 0x80) { // WE_HAVE_A_TWO_BY_TWO
            need(8);
            for (int i = 0; i < 4; i++) m[i] = (float)r.i16(p + 2 * (size_t)i) * k;
            p += 8;
        }
        // Every component visit counts against the budget too: a hostile font whose composites fan out into
        // empty glyphs would otherwise recurse ~1000^8 times without ever adding a point (FreeType caps the
        // same way through maxp.maxComponentElements / maxComponentDepth).
        total_points += MAX_GLYPH_POINTS / MAX_COMPONENT_VISITS;
        if (total_points > MAX_GLYPH_POINTS) throw Bad{};
        std::vector<GlyphContour> sub;
        read_contours(f, r, component, depth + 1, sub, total_points);
        for (GlyphContour &c : sub) {
            for (GlyphPoint &q : c) {
                const float x = q.x, y = q.y;
                q.x = m[0] * x + m[2] * y + dx;
                q.y = m[1] * x + m[3] * y + dy;
            }
            out.push_back(std::move(c));
        }
        if (!(cflags & 0x20)) break; // MORE_COMPONENTS
    }
}

// One TrueType contour -> pathfinder points: flag 0 = on the curve, CONTROL_POINT_0 = quadratic control point.
// The contour ends on a copy of its start point when its closing segment is a curve (pathfinder's implicit closing
// line is then empty); a straight closing segment is left to that implicit line.
void append_contour(const GlyphContour &c, PFOutline &out) {
    const size_t n = c.size();
    if (n == 0) return;
    PFVector2F start;
    size_t first = 0, count = n; // the points walked after the start point: c[first .. first + count)
    if (c[0].on_curve) {
        start = PFVector2F{c[0].x, c[0].y};
        first = 1, count = n - 1;
    } else if (c[n - 1].on_curve) {
        start = PFVector2F{c[n - 1].x, c[n - 1].y};
        first = 0, count = n - 1;
    } else {
        start = PFVector2F{(c[0].x + c[n - 1].x) * 0.5f, (c[0].y + c[n - 1].y) * 0.5f};
        first = 0, count = n;
    }
    const size_t base = out.points.size();
    auto push = [&](PFVector2F p, uint8_t flag) {
        out.points.push_back(p);
        out.flags.push_back(flag);
    };
    push(start, 0);
    bool have_control = false;
    PFVector2F control{0, 0};
    for (size_t i = 0; i < count; i++) {
        const GlyphPoint &q = c[first + i];
        const PFVector2F v{q.x, q.y};
        if (q.on_curve) {
            if (have_control) push(control, PF_POINT_FLAGS_CONTROL_POINT_0);
            have_control = false;
            push(v, 0);
        } else {
            if (have_control) {
                push(control, PF_POINT_FLAGS_CONTROL_POINT_0);
                push(PFVector2F{(control.x + v.x) * 0.5f, (control.y + v.y) * 0.5f}, 0);
            }
            control = v;
            have_control = true;
        }
    }
    if (have_control) {
        push(control, PF_POINT_FLAGS_CONTROL_POINT_0);
        push(start, 0);
    }
    if (out.points.size() - base < 3) { // a point or a single line encloses nothing
        out.points.resize(base);
        out.flags.resize(base);
        return;
    }
    out.contour_offsets.push_back((uint32_t)out.points.size());
    out.closed.push_back(1);
}

uint32_t lookup_format4(const Reader &r, uint32_t sub, uint32_t code) {
    if (code > 0xffff) return 0;
    const uint32_t seg_x2 = r.u16(sub + 6);
    const size_t ends = sub + 14, starts = ends + seg_x2 + 2, deltas = starts + seg_x2, offsets = deltas + seg_x2;
    for (uint32_t k = 0; k < seg_x2 / 2; k++) {
        const uint32_t end = r.u16(ends + 2 * (size_t)k);
        if (code > end) continue;
        const uint32_t start = r.u16(starts + 2 * (size_t)k);
        if (code < start) return 0;
        const uint16_t delta = r.u16(deltas + 2 * (size_t)k);
        const uint16_t range_offset = r.u16(offsets + 2 * (size_t)k);
        if (range_offset == 0) return (code + delta) & 0xffff;
        const uint16_t glyph = r.u16(offsets + 2 * (size_t)k + range_offset + 2 * (size_t)(code - start));
        return glyph ? (uint32_t)((glyph + delta) & 0xffff) : 0;
    }
    return 0;
}

uint32_t lookup_format12(const Reader &r, uint32_t sub, uint32_t code) {
    const uint32_t groups = r.u32(sub + 12);
    uint32_t lo = 0, hi = groups;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        const size_t g = (size_t)sub + 16 + 12 * (size_t)mid;
        const uint32_t start = r.u32(g), end = r.u32(g + 4);
        if (code < start) hi = mid;
        else if (code > end) lo = mid + 1;
        else return r.u32(g + 8) + (code - start);
    }
    return 0;
}

} // namespace

extern "C" {

PFFontRef PFFontCreateFromBytes(const uint8_t *data, size_t length) {
    if (!data || length < 12) {
        pf::set_last_error("PFFontCreateFromBytes: no data");
        return nullptr;
    }
    PFFont *f = new PFFont();
    try {
        f->data.assign(data, data + length);
        const Reader r{f->data};
        const uint32_t version = r.u32(0);
        if (version != 0x00010000u && version != 0x74727565u /* 'true' */) {
            pf::set_last_error("PFFontCreateFromBytes: not a TrueType-flavoured sfnt (CFF outlines and collections are not read)");
            delete f;
            return nullptr;
        }
        uint32_t head, head_len, maxp, maxp_len, hhea, hhea_len, cmap, cmap_len;
        if (!find_table(r, "head", head, head_len) || !find_table(r, "maxp", maxp, maxp_len) ||
            !find_table(r, "hhea", hhea, hhea_len) || !find_table(r, "hmtx", f->hmtx, f->hmtx_len) ||
            !find_table(r, "loca", f->loca, f->loca_len) || !find_table(r, "glyf", f->glyf, f->glyf_len) ||
            !find_table(r, "cmap", cmap, cmap_len) || head_len < 54 || maxp_len < 6 || hhea_len < 36)
            throw Bad{};
        f->units_per_em = r.u16(head + 18);
        f->loca_long = r.i16(head + 50) == 1;
        f->glyph_count = r.u16(maxp + 4);
        f->h_metrics = r.u16(hhea + 34);
        if (f->units_per_em == 0 || f->h_metrics == 0 || (uint64_t)4 * f->h_metrics > f->hmtx_len) throw Bad{};
        // Unicode subtables: (3,10) / (0,4+) full repertoire in format 12, (3,1) / (0,*) BMP in format 4.
        const uint16_t n = r.u16(cmap + 2);
        for (uint16_t i = 0; i < n; i++) {
            const size_t rec = (size_t)cmap + 4 + 8 * (size_t)i;
            const uint16_t platform = r.u16(rec), encoding = r.u16(rec + 2);
            const uint32_t off = r.u32(rec + 4);
            if (off > cmap_len) throw Bad{};
            const uint32_t sub = cmap + off;
            const uint16_t format = r.u16(sub);
            const bool unicode = platform == 0 || (platform == 3 && (encoding == 1 || encoding == 10));
            if (!unicode) continue;
            if (format == 4) {
                const uint32_t seg_x2 = r.u16(sub + 6);
                r.need(sub, 16 + 4 * (size_t)seg_x2);
                f->cmap4 = sub;
            } else if (format == 12) {
                const uint32_t groups = r.u32(sub + 12);
                r.need(sub, 16 + 12 * (size_t)groups);
                f->cmap12 = sub;
            }
        }
        if (!f->cmap4 && !f->cmap12) throw Bad{};
        return f;
    } catch (const Bad &) {
        pf::set_last_error("PFFontCreateFromBytes: truncated or malformed font tables");
    } catch (const std::exception &e) {
        pf::set_last_error(std::string("PFFontCreateFromBytes: ") + e.what());
    }
    delete f;
    return nullptr;
}

void PFFontDestroy(PFFontRef font) { delete font; }

uint32_t PFFontGetUnitsPerEm(PFFontRef font) { return font ? font->units_per_em : 0; }
uint32_t PFFontGetGlyphCount(PFFontRef font) { return font ? font->glyph_count : 0; }

uint32_t PFFontGetGlyphForCodepoint(PFFontRef font, uint32_t codepoint) {
    if (!font) return 0;
    try {
        const Reader r{font->data};
        uint32_t glyph = 0;
        if (font->cmap12) glyph = lookup_format12(r, font->cmap12, codepoint);
        if (!glyph && font->cmap4) glyph = lookup_format4(r, font->cmap4, codepoint);
        return glyph < font->glyph_count ? glyph : 0;
    } catch (const Bad &) {
        return 0;
    }
}

float PFFontGetGlyphAdvance(PFFontRef font, uint32_t glyph_id) {
    if (!font || glyph_id >= font->glyph_count) return 0.0f;
    try {
        const Reader r{font->data};
        const uint32_t k = glyph_id < font->h_metrics ? glyph_id : font->h_metrics - 1; // trailing glyphs share the last advance
        return (float)r.u16((size_t)font->hmtx + 4 * (size_t)k);
    } catch (const Bad &) {
        return 0.0f;
    }
}

PFOutlineRef PFFontGetGlyphOutline(PFFontRef font, uint32_t glyph_id) {
    if (!font) {
        pf::set_last_error("PFFontGetGlyphOutline: null font");
        return nullptr;
    }
    PFOutline *outline = new PFOutline();
    try {
        const Reader r{font->data};
        std::vector<GlyphContour> contours;
        size_t total_points = 0;
        read_contours(*font, r, glyph_id, 0, contours, total_points);
        for (const GlyphContour &c : contours) append_contour(c, *outline);
        return outline;
    } catch (const Bad &) {
        pf::set_last_error("PFFontGetGlyphOutline: glyph id out of range or malformed glyph data");
    } catch (const std::exception &e) {
        pf::set_last_error(std::string("PFFontGetGlyphOutline: ") + e.what());
    }
    delete outline;
    return nullptr;
}

} // extern "C"
