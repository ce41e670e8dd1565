// tools/fuzz/font_asan.cpp — truncated / corrupted font files against csrc/font.cpp under ASan + UBSan (see run.sh).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../include/pf_cuda.h"
namespace pf { void set_last_error(const std::string &) {} }
extern "C" {
uint32_t PFOutlineGetContourCount(PFOutlineRef);
}
#include "../../pathfinder_b200/csrc/outline.h"
extern "C" void PFOutlineDestroy(PFOutlineRef o) { delete o; }
static void poke(const std::vector<uint8_t> &blob) {
    PFFontRef h = PFFontCreateFromBytes(blob.data(), blob.size());
    if (!h) return;
    uint32_t n = PFFontGetGlyphCount(h);
    for (uint32_t g = 0; g < n + 2 && g < 1500; g++) {
        PFFontGetGlyphAdvance(h, g);
        PFOutlineRef o = PFFontGetGlyphOutline(h, g);
        if (o) PFOutlineDestroy(o);
    }
    for (uint32_t c = 0; c < 0x3000; c += 7) PFFontGetGlyphForCodepoint(h, c);
    PFFontGetGlyphForCodepoint(h, 0x1F600);
    PFFontDestroy(h);
}
int main(int argc, char **argv) {
    FILE *f = fopen(argc > 1 ? argv[1] : "/root/reference/resources/fonts/Roboto-Regular.ttf", "rb");
    if (!f) { puts("no font file: pass a .ttf path"); return 0; }
    std::vector<uint8_t> data(1 << 20);
    data.resize(fread(data.data(), 1, data.size(), f));
    fclose(f);
    poke(data);
    for (size_t n = 0; n < data.size(); n += (n < 2000 ? 3 : 4099)) poke(std::vector<uint8_t>(data.begin(), data.begin() + n));
    srand(7);
    for (int it = 0; it < 400; it++) {
        std::vector<uint8_t> blob = data;
        size_t lo = it % 3 == 0 ? 0 : 0, hi = it % 3 == 0 ? 400 : (it % 3 == 1 ? 6000 : blob.size());
        for (int k = 0; k < (it % 3 == 2 ? 300 : 10); k++) blob[lo + rand() % (hi - lo)] ^= 1 + rand() % 255;
        poke(blob);
    }
    puts("ok");
}
