// fft_core.cuh — mixed-radix Stockham FFT building blocks for the FDK ramp filter.
//
// The reference convolves every detector row directly (recon/bp3d20.cpp:63-73, O(nu^2) per row).  For
// wide detectors the same linear convolution is done as zero-padded circular convolution of length
// L >= 2 nu: forward FFT, product with the (real, even) spectrum of the taps, inverse FFT — two real
// rows ride in one complex transform.  One butterfly per thread and pass, passes separated by a
// barrier, data exchanged through two padded shared-memory buffers (auto-sort: no bit reversal).
//
// Everything here is plain C++ so that the same code is driven by CUDA threads in fdk.cu and by a
// sequential loop over "threads" in tests/fft_core_host.cpp (CPU test of the index algebra).
#pragma once

#ifdef __CUDACC__
#define MONTE_HD __host__ __device__ __forceinline__
#else
#define MONTE_HD inline
#endif

namespace monte {

struct alignas(8) cpx { float x, y; };

MONTE_HD cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
MONTE_HD cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
MONTE_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
// a * (S i),  S = -1 forward, +1 inverse
template <int S> MONTE_HD cpx cmuli(cpx a) { return S > 0 ? cpx{-a.y, a.x} : cpx{a.y, -a.x}; }

// X[q] = sum_r v[r] exp(S 2 pi i q r / R), in place, natural order
template <int S> MONTE_HD void dft2(cpx *v) {
    const cpx a = v[0];
    v[0] = cadd(a, v[1]); v[1] = csub(a, v[1]);
}
template <int S> MONTE_HD void dft4(cpx &a0, cpx &a1, cpx &a2, cpx &a3) {
    const cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmuli<S>(csub(a1, a3));
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}
template <int S> MONTE_HD void dft8(cpx *v) {
    dft4<S>(v[0], v[2], v[4], v[6]);                 // even inputs -> E[0..3] in v[0],v[2],v[4],v[6]
    dft4<S>(v[1], v[3], v[5], v[7]);                 // odd inputs  -> O[0..3] in v[1],v[3],v[5],v[7]
    const float h = 0.70710678118654752f, s = (float)S;
    const cpx o0 = v[1];
    const cpx o1 = cpx{(v[3].x - s * v[3].y) * h, (s * v[3].x + v[3].y) * h};      // * exp(S i pi/4)
    const cpx o2 = cmuli<S>(v[5]);
    const cpx o3 = cpx{(-v[7].x - s * v[7].y) * h, (s * v[7].x - v[7].y) * h};     // * exp(S 3 i pi/4)
    const cpx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, o0); v[1] = cadd(e1, o1); v[2] = cadd(e2, o2); v[3] = cadd(e3, o3);
    v[4] = csub(e0, o0); v[5] = csub(e1, o1); v[6] = csub(e2, o2); v[7] = csub(e3, o3);
}
template <int S, int R> MONTE_HD void dftR(cpx *v) {
    if (R == 8) dft8<S>(v);
    else if (R == 4) dft4<S>(v[0], v[1], v[2], v[3]);
    else dft2<S>(v);
}

// exchange-buffer index: XOR swizzle of the low four bits (an 8-byte element occupies one of 16 bank
// pairs; a half-warp is served per wavefront).  With m = i >> 4 the swizzle (m & 7) | ((m & 4) << 1)
// keeps aligned runs of 16 a permutation of the 16 bank pairs (all loads, stores of the Ns >= 64
// passes), sends the stride-8 stores of the first pass to 16 distinct pairs (low three bits of m
// distinct over 8 consecutive blocks) and separates the two 8-runs, 64 elements apart, that a
// half-warp stores in the Ns = 8 pass (bit 3 follows bit 2 of m).  No padding is needed.
MONTE_HD int fft_pad(int i) { const int m = i >> 4; return i ^ ((m & 7) | ((m & 4) << 1)); }
constexpr int fft_padded_len(int L) { return L; }

// radix plan: three radix-8 passes (Ns = 1, 8, 64) and a last pass of radix L/512 at Ns = 512
template <int L> struct FftPlan {
    static_assert(L == 1024 || L == 2048 || L == 4096, "supported transform lengths");
    static constexpr int THREADS = L / 8;
    static constexpr int R_LAST = L / 512;            // 2, 4 or 8
    static constexpr int B_LAST = 8 / R_LAST;         // butterflies per thread in the last pass
};

// One radix-R butterfly of the Stockham pass with sub-transform length Ns (product of the radices of
// the passes before it): reads element jj + r L/R, r = 0..R-1, writes (jj / Ns) Ns R + jj % Ns + q Ns.
// w1 = exp(-2 pi i (jj % Ns) / (Ns R)) is fixed per thread and pass, so it lives in a register and its
// powers are formed by multiplication (w2 = w1^2, w3 = w2 w1, w4 = w2^2, w5 = w4 w1, w6 = w4 w2,
// w7 = w4 w3: at most three roundings) instead of seven table fetches through the shared-memory pipe,
// which is what bounds the kernel.  The inverse transform uses the conjugate.
template <int L, int S, int R, int NS, class Load, class Store>
MONTE_HD void fft_butterfly(int jj, cpx w1, Load ld, Store st) {
    constexpr int STRIDE = L / R;
    const int k = jj & (NS - 1);
    cpx v[R];
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = ld(jj + r * STRIDE);
    if (NS > 1) {
        if (S > 0) w1.y = -w1.y;
        v[1] = cmul(v[1], w1);
        if (R >= 4) {
            const cpx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
            if (R == 8) {
                const cpx w4 = cmul(w2, w2);
                v[4] = cmul(v[4], w4); v[5] = cmul(v[5], cmul(w4, w1));
                v[6] = cmul(v[6], cmul(w4, w2)); v[7] = cmul(v[7], cmul(w4, w3));
            }
        }
    }
    dftR<S, R>(v);
    const int j0 = (jj - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; q++) st(j0 + q * NS, v[q]);
}

// per-thread twiddles of the three twiddled passes (Ns = 8, 64, 512), read once from
// tw[m] = exp(-2 pi i m / L)
template <int L> struct FftTwiddles {
    cpx p2, p3, p4[FftPlan<L>::B_LAST];
    MONTE_HD void load(const cpx *tw, int j) {
        using P = FftPlan<L>;
        p2 = tw[(j & 7) * (L / 64)];
        p3 = tw[(j & 63) * (L / 512)];
#pragma unroll
        for (int m = 0; m < P::B_LAST; m++) p4[m] = tw[((j + m * P::THREADS) & 511) * (L / (512 * P::R_LAST))];
    }
};

// The filter of one row pair as seen by "thread" j of FftPlan<L>::THREADS; `phase` 0..7 are the
// eight passes, each followed by a barrier in the kernel.  in(n) yields the zero-padded complex input
// (row A + i row B), spec[n] the real spectrum of the taps divided by L, out(n, value) takes the
// filtered pair for n < L (callers keep n < nu).
template <int L, class In, class Out>
MONTE_HD void fft_filter_phase(int phase, int j, cpx *bufA, cpx *bufB, const FftTwiddles<L> &w, const float *spec, In in, Out out) {
    using P = FftPlan<L>;
    auto ldA = [&](int i) { return bufA[fft_pad(i)]; };
    auto ldB = [&](int i) { return bufB[fft_pad(i)]; };
    auto stA = [&](int i, cpx v) { bufA[fft_pad(i)] = v; };
    auto stB = [&](int i, cpx v) { bufB[fft_pad(i)] = v; };
    const cpx one = cpx{1.f, 0.f};
    switch (phase) {
    case 0: fft_butterfly<L, -1, 8, 1>(j, one, in, stA); break;
    case 1: fft_butterfly<L, -1, 8, 8>(j, w.p2, ldA, stB); break;
    case 2: fft_butterfly<L, -1, 8, 64>(j, w.p3, ldB, stA); break;
    case 3:
#pragma unroll
        for (int m = 0; m < P::B_LAST; m++) fft_butterfly<L, -1, P::R_LAST, 512>(j + m * P::THREADS, w.p4[m], ldA, stB);
        break;
    case 4:
        fft_butterfly<L, +1, 8, 1>(j, one, [&](int i) { const cpx a = bufB[fft_pad(i)]; const float h = spec[i]; return cpx{a.x * h, a.y * h}; }, stA);
        break;
    case 5: fft_butterfly<L, +1, 8, 8>(j, w.p2, ldA, stB); break;
    case 6: fft_butterfly<L, +1, 8, 64>(j, w.p3, ldB, stA); break;
    default:
#pragma unroll
        for (int m = 0; m < P::B_LAST; m++) fft_butterfly<L, +1, P::R_LAST, 512>(j + m * P::THREADS, w.p4[m], ldA, out);
        break;
    }
}

}  // namespace monte
