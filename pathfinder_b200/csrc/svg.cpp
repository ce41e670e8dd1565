// pathfinder_b200/csrc/svg.cpp — SVG path data -> outline (front-end half). The reference gets
// its outlines from usvg 0.9.1 (svg/src/lib.rs:386-456 walks usvg's normalised MoveTo / LineTo / CurveTo /
// ClosePath segments); usvg is not vendored, so this follows the published rules instead (W3C SVG 1.1 §8.3 path
// grammar, implicit commands, smooth-curve reflection, §F.6 arc conversion: one cubic per <= 90 degree piece) in
// double precision like usvg, and converts to f32 points at the end (svg/src/lib.rs:441-453). Parity: unpinned
// (no reference test pins usvg); cross-checked against the Python parser that built the tiger fixture.
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pf_cuda.h"
#include "outline.h"

namespace pf {
void set_last_error(const std::string &msg);
}

namespace {

struct P {
    double x, y;
};

struct Builder {
    PFOutline *out;
    bool open = false;
    void begin(P p) {
        end(false);
        push(p, 0);
        open = true;
    }
    void push(P p, uint8_t flag) {
        out->points.push_back(PFVector2F{(float)p.x, (float)p.y});
        out->flags.push_back(flag);
    }
    void end(bool closed) {
        if (!open) return;
        open = false;
        if (out->points.size() - out->contour_offsets.back() < 2) { // a lone moveto draws nothing
            out->points.resize(out->contour_offsets.back());
            out->flags.resize(out->contour_offsets.back());
            return;
        }
        out->contour_offsets.push_back((uint32_t)out->points.size());
        out->closed.push_back(closed ? 1 : 0);
    }
};

struct Parser {
    const char *s;
    size_t pos = 0, n;
    bool failed = false;
    explicit Parser(const char *d) : s(d), n(strlen(d)) {}
    void skip() {
        while (pos < n && (isspace((unsigned char)s[pos]) || s[pos] == ',')) pos++;
    }
    double num() {
        skip();
        char *end = nullptr;
        const double v = strtod(s + pos, &end);
        if (end == s + pos) {
            failed = true;
            return 0.0;
        }
        pos = (size_t)(end - s);
        return v;
    }
    bool flag() {
        skip();
        if (pos >= n || (s[pos] != '0' && s[pos] != '1')) {
            failed = true;
            return false;
        }
        return s[pos++] == '1';
    }
};

// SVG 1.1 F.6.5 / F.6.6: endpoint -> centre parameterisation, then one cubic per <= 90 degree piece.
void arc_to_cubics(Builder &b, P p0, double rx, double ry, double phi_deg, bool large, bool sweep, P p1) {
    if (p0.x == p1.x && p0.y == p1.y) return;
    rx = std::fabs(rx), ry = std::fabs(ry);
    if (rx == 0.0 || ry == 0.0) {
        b.push(p1, 0);
        return;
    }
    const double pi = 3.14159265358979323846;
    const double phi = phi_deg * pi / 180.0, cphi = std::cos(phi), sphi = std::sin(phi);
    const double dx2 = (p0.x - p1.x) / 2.0, dy2 = (p0.y - p1.y) / 2.0;
    const double x1p = cphi * dx2 + sphi * dy2, y1p = -sphi * dx2 + cphi * dy2;
    const double lam = (x1p * x1p) / (rx * rx) + (y1p * y1p) / (ry * ry);
    const bool scaled_up = lam > 1.0; // F.6.6.3: radii too small for the chord are scaled until it is a diameter
    if (scaled_up) rx *= std::sqrt(lam), ry *= std::sqrt(lam);
    const double num = rx * rx * ry * ry - rx * rx * y1p * y1p - ry * ry * x1p * x1p;
    const double den = rx * rx * y1p * y1p + ry * ry * x1p * x1p;
    // With scaled-up radii the centre is the chord's midpoint exactly; computing it from num / den leaves a
    // rounding residue of either sign there, which makes the sweep (exactly half a turn) come out as pi +- 1e-8 and
    // the number of pieces depend on the compiler.
    const double coef = scaled_up ? 0.0 : std::sqrt(std::fmax(num / den, 0.0)) * (large == sweep ? -1.0 : 1.0);
    const double cxp = coef * rx * y1p / ry, cyp = -coef * ry * x1p / rx;
    const double cx = cphi * cxp - sphi * cyp + (p0.x + p1.x) / 2.0, cy = sphi * cxp + cphi * cyp + (p0.y + p1.y) / 2.0;
    auto angle = [](double ux, double uy, double vx, double vy) { return std::atan2(ux * vy - uy * vx, ux * vx + uy * vy); };
    const double th1 = angle(1.0, 0.0, (x1p - cxp) / rx, (y1p - cyp) / ry);
    double dth = angle((x1p - cxp) / rx, (y1p - cyp) / ry, (-x1p - cxp) / rx, (-y1p - cyp) / ry);
    if (!sweep && dth > 0.0)
        dth -= 2.0 * pi;
    else if (sweep && dth < 0.0)
        dth += 2.0 * pi;
    const int pieces = std::max(1, (int)std::ceil(std::fabs(dth) / (pi / 2.0) - 1e-9));
    const double delta = dth / pieces, k = 4.0 / 3.0 * std::tan(delta / 4.0);
    auto point = [&](double th) {
        const double x = rx * std::cos(th), y = ry * std::sin(th);
        return P{cphi * x - sphi * y + cx, sphi * x + cphi * y + cy};
    };
    auto derivative = [&](double th) {
        const double x = -rx * std::sin(th), y = ry * std::cos(th);
        return P{cphi * x - sphi * y, sphi * x + cphi * y};
    };
    P cur = p0;
    for (int i = 0; i < pieces; i++) {
        const double a0 = th1 + i * delta, a1 = th1 + (i + 1) * delta;
        const P e = i == pieces - 1 ? p1 : point(a1), d0 = derivative(a0), d1 = derivative(a1);
        b.push(P{cur.x + k * d0.x, cur.y + k * d0.y}, PF_POINT_FLAGS_CONTROL_POINT_0);
        b.push(P{e.x - k * d1.x, e.y - k * d1.y}, PF_POINT_FLAGS_CONTROL_POINT_1);
        b.push(e, 0);
        cur = e;
    }
}

} // namespace

extern "C" {

PFOutlineRef PFSvgPathDataToOutline(const char *d) {
    if (!d) {
        pf::set_last_error("PFSvgPathDataToOutline: null path data");
        return nullptr;
    }
    PFOutline *out = new PFOutline;
    Builder b{out};
    Parser p(d);
    P cur{0, 0}, start{0, 0}, last_ctrl{0, 0};
    char cmd = 0, last = 0;
    bool have_ctrl = false;
    for (;;) {
        p.skip();
        if (p.pos >= p.n) break;
        if (isalpha((unsigned char)p.s[p.pos])) {
            cmd = p.s[p.pos++];
            if (cmd == 'Z' || cmd == 'z') {
                b.end(true);
                cur = start;
                last = 'Z';
                continue;
            }
        } else if (cmd == 'M') {
            cmd = 'L'; // implicit lineto after moveto
        } else if (cmd == 'm') {
            cmd = 'l';
        } else if (cmd == 0) {
            p.failed = true;
        }
        if (p.failed) break;
        const bool rel = islower((unsigned char)cmd) != 0;
        const char c = (char)toupper((unsigned char)cmd);
        const P o = rel ? cur : P{0, 0};
        if (c == 'M') {
            const double x = o.x + p.num(), y = o.y + p.num();
            cur = start = P{x, y};
            b.begin(cur);
        } else {
            if (!b.open) { // drawing after Z without M: a new subpath at the old start
                start = cur;
                b.begin(cur);
            }
            if (c == 'L') {
                const double x = o.x + p.num(), y = o.y + p.num();
                cur = P{x, y};
                b.push(cur, 0);
            } else if (c == 'H') {
                cur = P{(rel ? cur.x : 0.0) + p.num(), cur.y};
                b.push(cur, 0);
            } else if (c == 'V') {
                cur = P{cur.x, (rel ? cur.y : 0.0) + p.num()};
                b.push(cur, 0);
            } else if (c == 'C' || c == 'S') {
                P c0 = cur;
                if (c == 'C') {
                    const double x = o.x + p.num(), y = o.y + p.num();
                    c0 = P{x, y};
                } else if ((last == 'C' || last == 'S') && have_ctrl) {
                    c0 = P{2 * cur.x - last_ctrl.x, 2 * cur.y - last_ctrl.y};
                }
                const double x1 = o.x + p.num(), y1 = o.y + p.num();
                const double x = o.x + p.num(), y = o.y + p.num();
                b.push(c0, PF_POINT_FLAGS_CONTROL_POINT_0);
                b.push(P{x1, y1}, PF_POINT_FLAGS_CONTROL_POINT_1);
                last_ctrl = P{x1, y1}, have_ctrl = true;
                cur = P{x, y};
                b.push(cur, 0);
            } else if (c == 'Q' || c == 'T') {
                P q = cur;
                if (c == 'Q') {
                    const double x = o.x + p.num(), y = o.y + p.num();
                    q = P{x, y};
                } else if ((last == 'Q' || last == 'T') && have_ctrl) {
                    q = P{2 * cur.x - last_ctrl.x, 2 * cur.y - last_ctrl.y};
                }
                const double x = o.x + p.num(), y = o.y + p.num();
                b.push(q, PF_POINT_FLAGS_CONTROL_POINT_0);
                last_ctrl = q, have_ctrl = true;
                cur = P{x, y};
                b.push(cur, 0);
            } else if (c == 'A') {
                const double rx = p.num(), ry = p.num(), rotation = p.num();
                const bool large = p.flag(), sweep = p.flag();
                const double x = o.x + p.num(), y = o.y + p.num();
                if (!p.failed) arc_to_cubics(b, cur, rx, ry, rotation, large, sweep, P{x, y});
                cur = P{x, y};
            } else {
                p.failed = true;
            }
        }
        if (p.failed) break;
        last = c;
    }
    if (p.failed) {
        pf::set_last_error("PFSvgPathDataToOutline: malformed path data near offset " + std::to_string(p.pos));
        delete out;
        return nullptr;
    }
    b.end(false);
    return out;
}

} // extern "C"
