// tools/fuzz/host_asan.cpp — random and grammar-aware SVG path data through csrc/svg.cpp, every stroke style of csrc/stroke.cpp and
// csrc/dilate.cpp under ASan + UBSan; the input being processed is kept in last_case.txt (see run.sh).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include "../../include/pf_cuda.h"
namespace pf { void set_last_error(const std::string &) {} }
int main() {
    srand(11);
    const char alphabet[] = "MmLlHhVvCcSsQqTtAaZz0123456789.,-+eE \t\n";
    int parsed = 0;
    for (int it = 0; it < 200000; it++) {
        std::string s;
        if (it % 2 == 0) {
            int len = rand() % 60;
            if (it % 4 == 0) s = "M10 10";
            for (int i = 0; i < len; i++) s += alphabet[rand() % (sizeof(alphabet) - 1)];
        } else {
            static const char cmds[] = "MmLlHhVvCcSsQqTtAaZz";
            static const int argc[] = {2,2,2,2,1,1,1,1,6,6,4,4,4,4,2,2,7,7,0,0};
            s = "M";
            s += std::to_string(rand() % 200 - 50) + " " + std::to_string(rand() % 200 - 50);
            int n = rand() % 12;
            for (int c = 0; c < n; c++) {
                int ci = rand() % 20;
                s += ' ';
                s += cmds[ci];
                int reps = 1 + rand() % 2;
                for (int a = 0; a < argc[ci] * reps + (rand() % 16 == 0); a++) {
                    char buf[64];
                    double v = (rand() % 40000 - 20000) / (double)(1 + rand() % 300);
                    if ((ci == 16 || ci == 17) && (a % 7 == 3 || a % 7 == 4)) v = rand() % 2;
                    if (rand() % 50 == 0) v = 0;
                    if (rand() % 200 == 0) v = 1e30;
                    snprintf(buf, sizeof buf, rand() % 8 ? "%g" : "%.3e", v);
                    s += (a || rand() % 2) ? " " : "";
                    s += buf;
                    if (rand() % 3 == 0) s += ",";
                }
            }
        }
        { FILE *lf = fopen("last_case.txt", "w"); fprintf(lf, "%d parse: %s\n", it, s.c_str()); fclose(lf); }
        PFOutlineRef o = PFSvgPathDataToOutline(s.c_str());
        if (!o) continue;
        parsed++;
        size_t n = PFOutlineGetPointCount(o);
        uint32_t k = PFOutlineGetContourCount(o);
        std::vector<PFVector2F> pts(n + 1);
        std::vector<uint8_t> fl(n + 1), closed(k + 1);
        std::vector<uint32_t> off(k + 1);
        PFOutlineCopy(o, pts.data(), fl.data(), off.data());
        PFOutlineCopyClosed(o, closed.data());
        // stroke whatever came out, with every style, then dilate it
        for (int style = 0; style < 9 && n; style++) {
            PFStrokeStyle st{(float)(rand() % 100) / 7.0f, (uint32_t)(style % 3), (uint32_t)(style / 3), (float)(rand() % 20)};
            { FILE *lf = fopen("last_case.txt", "a"); fprintf(lf, "stroke style %d width %g miter %g\n", style, st.line_width, st.miter_limit); fclose(lf); }
            PFOutlineRef so = PFOutlineStrokeToFill(pts.data(), fl.data(), off.data(), closed.data(), k, &st);
            if (so) {
                size_t sn = PFOutlineGetPointCount(so);
                uint32_t sk = PFOutlineGetContourCount(so);
                std::vector<PFVector2F> sp(sn + 1);
                std::vector<uint8_t> sf(sn + 1);
                std::vector<uint32_t> sof(sk + 1);
                PFOutlineCopy(so, sp.data(), sf.data(), sof.data());
                PFVector2F amount{0.3f, 0.7f};
                PFOutlineDilate(sp.data(), sof.data(), sk, &amount);
                PFOutlineDestroy(so);
            }
        }
        PFVector2F amount{1.0f, 0.5f};
        PFOutlineDilate(pts.data(), off.data(), k, &amount);
        PFOutlineDestroy(o);
    }
    printf("ok, %d parsed\n", parsed);
}
