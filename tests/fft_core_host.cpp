// CPU check of monte_b200/csrc/fft_core.cuh: the per-thread FFT phases, driven sequentially, must
// reproduce the direct Ram-Lak convolution (recon/bp3d20.cpp:63-73) of two rows.
// Build: g++ -O2 -std=c++17 -I monte_b200/csrc tests/fft_core_host.cpp -o <exe>; prints one line per case.
#include "fft_core.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace monte;

template <int L> static double run_case(int nu, unsigned seed) {
    using P = FftPlan<L>;
    std::vector<float> a(nu), b(nu);
    srand(seed);
    for (int i = 0; i < nu; i++) { a[i] = rand() / (float)RAND_MAX; b[i] = rand() / (float)RAND_MAX - 0.3f; }
    std::vector<double> g(L, 0.0);
    g[0] = (float)(0.5 * 0.25);
    for (int n = 1; n < nu; n++)
        if (n & 1) { float ramp = (float)(-1. / pow(n * M_PI, 2)); g[n] = g[L - n] = (float)(0.5 * (double)ramp); }
    std::vector<float> spec(L);
    for (int k = 0; k < L; k++) {
        double s = 0;
        for (int n = 0; n < L; n++) s += g[n] * cos(2 * M_PI * (double)((long long)k * n % L) / L);
        spec[k] = (float)(s / L);
    }
    std::vector<cpx> tw(L), A(fft_padded_len(L)), B(fft_padded_len(L));
    for (int m = 0; m < L; m++) tw[m] = cpx{(float)cos(-2 * M_PI * m / L), (float)sin(-2 * M_PI * m / L)};
    std::vector<float> oa(nu, 0.f), ob(nu, 0.f);
    auto in = [&](int n) { return n < nu ? cpx{a[n], b[n]} : cpx{0.f, 0.f}; };
    auto out = [&](int n, cpx v) { if (n < nu) { oa[n] = v.x; ob[n] = v.y; } };
    for (int phase = 0; phase < 8; phase++)
        for (int j = 0; j < P::THREADS; j++) {
            FftTwiddles<L> w;
            w.load(tw.data(), j);
            fft_filter_phase<L>(phase, j, A.data(), B.data(), w, spec.data(), in, out);
        }
    double err = 0, mx = 0;
    for (int o = 0; o < nu; o++) {
        double da = 0, db = 0;
        for (int c = 0; c < nu; c++) { const double t = g[((c - o) % L + L) % L]; da += a[c] * t; db += b[c] * t; }
        err = fmax(err, fmax(fabs(da - oa[o]), fabs(db - ob[o])));
        mx = fmax(mx, fmax(fabs(da), fabs(db)));
    }
    return err / mx;
}

int main() {
    int bad = 0;
    struct { int L, nu; } cases[] = {{1024, 512}, {1024, 325}, {2048, 1024}, {2048, 700}, {4096, 2048}, {4096, 1536}};
    for (auto c : cases) {
        double e = c.L == 1024 ? run_case<1024>(c.nu, 1) : c.L == 2048 ? run_case<2048>(c.nu, 2) : run_case<4096>(c.nu, 3);
        printf("L=%d nu=%d rel_err=%.3e\n", c.L, c.nu, e);
        if (!(e < 2e-6)) bad++;
    }
    return bad;
}
