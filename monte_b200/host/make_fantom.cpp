// make_fantom — label-phantom generator (driver surface of monte_cpp/make_fantom.cpp, make_image01.cpp).
//   make_fantom disc   [out=ball_fan.raw]                 65x65 uint8 disc r=10 at (32,46)   make_fantom.cpp:10-19
//   make_fantom sphere [out=spher01.raw]                  185x185x325 sphere r=50 at (90,90,160) make_image01.cpp:15-23
//   make_fantom cylinder N pitch_cm [out=cyl.raw]         N^3 water cylinder r=10 cm + 8 Ca rods (CBCT_real325.cu:916-921)
// Files are headerless uint8, x fastest — the reference's layout.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "monte_gpu.h"

static int write_raw(const std::string &fn, const void *p, size_t bytes) {
    FILE *f = fopen(fn.c_str(), "wb");
    if (!f) { fprintf(stderr, "failed to open %s\n", fn.c_str()); return -1; }
    size_t w = fwrite(p, 1, bytes, f);
    fclose(f);
    if (w != bytes) { fprintf(stderr, "failed to write %s\n", fn.c_str()); return -1; }
    return 0;
}

int main(int argc, char **argv) {
    const std::string kind = argc > 1 ? argv[1] : "disc";
    if (kind == "disc") {
        std::vector<uint8_t> g(65 * 65);
        monte_make_fantom(g.data(), 65, 32, 46, 100);
        return write_raw(argc > 2 ? argv[2] : "ball_fan.raw", g.data(), g.size());
    }
    if (kind == "sphere") {
        std::vector<uint8_t> g((size_t)185 * 185 * 325);
        monte_make_sphere(g.data(), 185, 185, 325, 90, 90, 160, 2500);
        return write_raw(argc > 2 ? argv[2] : "spher01.raw", g.data(), g.size());
    }
    if (kind == "cylinder" && argc >= 4) {
        const int n = atoi(argv[2]);
        const double pitch = atof(argv[3]);
        std::vector<uint8_t> g((size_t)n * n * n, 0);
        for (int k = 0; k < n; k++)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++) {
                    const double x = (i + 0.5) * pitch - 0.5 * n * pitch, y = (j + 0.5) * pitch - 0.5 * n * pitch,
                                 z = (k + 0.5) * pitch - 0.5 * n * pitch;
                    uint8_t l = (x * x + y * y <= 100.0 && fabs(z) <= 10.0) ? 1 : 0;
                    if (l) for (int a = 0; a < 8; a++) {
                        const double cx = 5.0 * cos(a * M_PI / 4), cy = 5.0 * sin(a * M_PI / 4);
                        if ((x - cx) * (x - cx) + (y - cy) * (y - cy) <= 2.25) l = 2;
                    }
                    g[((size_t)k * n + j) * n + i] = l;
                }
        return write_raw(argc > 4 ? argv[4] : "cyl.raw", g.data(), g.size());
    }
    fprintf(stderr, "usage: make_fantom disc|sphere [out] | cylinder N pitch [out]\n");
    return 2;
}
