// Exercises the C++ host mirror (include/dplug_b200/compositor.hpp) against libb2pbr.so.
// Without a GPU: checks that construction fails loudly (B2_ERR_NO_DEVICE) and that the host-side tile scheduler
// works.  With a GPU (argv[1] == "gpu"): composites a small frame through ICompositor::compositeTile and prints
// a checksum of the result so the Python test can compare it with the ctypes path.
#include <cstdio>
#include <cstring>
#include <vector>
#include "dplug_b200/compositor.hpp"

int main(int argc, char** argv) {
    std::vector<b2::box2i> tiles = b2::tileAreas({b2::box2i{0, 0, 620, 330}}, 64, 64);
    if (tiles.size() != 60) { printf("tileAreas: %zu tiles\n", tiles.size()); return 1; }
    {   // DirtyRectList (boxlist.d:200-303): host logic, works without a GPU
        b2::DirtyRectList dirty;
        dirty.addRect(b2::box2i{0, 0, 10, 10});
        dirty.addRect(b2::box2i{5, 5, 15, 15});
        std::vector<b2::box2i> pulled;
        dirty.pullAllRectangles(pulled);
        if (pulled.size() != 3 || !b2::haveNoOverlap(pulled) || !dirty.isEmpty()) { printf("DirtyRectList: %zu rects\n", pulled.size()); return 1; }
    }
    const bool gpu = argc > 1 && !strcmp(argv[1], "gpu");
    if (!gpu) {
        if (b2_device_count() > 0) { printf("skip: a GPU is present\n"); return 0; }
        try {
            b2::PBRCompositor c(0);
        } catch (const b2::Error& e) {
            if (e.code == B2_ERR_NO_DEVICE) { printf("ok: no device -> %s\n", e.what()); return 0; }
        }
        printf("expected B2_ERR_NO_DEVICE\n");
        return 1;
    }
    const int W = 160, H = 96;
    std::vector<b2::RGBA> d(W * H), m(W * H), out(W * H);
    std::vector<b2::L16> z(W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            d[y * W + x] = b2::RGBA{(unsigned char)(x * 3 + y), (unsigned char)(x + y * 2), (unsigned char)(x ^ y), (unsigned char)((x / 16 + y / 16) % 7 == 0 ? 255 : 0)};
            m[y * W + x] = b2::RGBA{128, 64, 128, 255};
            z[y * W + x] = b2::L16{(unsigned short)(15000 + (x * 37 + y * 91) % 3000)};
        }
    b2::PBRCompositor comp(0);
    b2::ICompositor& ic = comp;
    ic.resizeBuffers(W, H, 64, 64);
    ic.compositeTile(b2::ImageRef<b2::RGBA>{W, H, (size_t)W * 4, out.data()}, b2::tileAreas({b2::box2i{0, 0, W, H}}, 64, 64),
                     b2::ImageRef<b2::RGBA>{W, H, (size_t)W * 4, d.data()}, b2::ImageRef<b2::RGBA>{W, H, (size_t)W * 4, m.data()},
                     b2::ImageRef<b2::L16>{W, H, (size_t)W * 2, z.data()});
    unsigned long long sum = 0;
    for (int k = 0; k < W * H; ++k) sum = sum * 1000003ull + out[k].r + 256u * out[k].g + 65536u * out[k].b + 16777216u * out[k].a;
    printf("checksum %llu\n", sum);
    return 0;
}
