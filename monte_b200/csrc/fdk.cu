// fdk.cu — FDK reconstruction kernels for sm_100a and their C-ABI entry points.
//
// Replaces recon/bp3d20.cpp:36-166 (and bp3d20_325.cpp, fbp2.cpp) of the reference:
//   fdk_weight_filter_kernel  = step 1 + step 2 (bp3d20.cpp:36-43, 48-73): cosine weight fused into
//                               the load of a shared-memory Ram-Lak convolution; writes transposed.
//   fdk_pad_kernel            = materialises the reference's "one past the row/view" reads
//                               (bp3d20.cpp:152-156) as a duplicated column + two zero rows.
//   fdk_backproject_kernel    = step 3 (bp3d20.cpp:83-166): voxel-driven, one thread per (s,t)
//                               column of ZT z-slices, all per-(s,t,view) terms hoisted, bilinear in
//                               registers, fp32 accumulation over views in registers.
//   fdk_transpose_kernel      = image_zy (bp3d20.cpp:161) as a tiled transpose of image_xy.
//   fbp2_* kernels            = recon/fbp2.cpp in double (nearest-neighbour lookup is discontinuous,
//                               fp32 coordinates would flip pixels).
// Neither kernel is a contraction worth tensor cores: the filter is a 1-D convolution with
// per-row taps, the backprojector a gather.  Bound: FP32 pipe + L1 gather.
#include "common.cuh"
#include "fft_core.cuh"
#ifndef MONTE_EMU
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <string>
#include <thread>

namespace monte {

// ------------------------------------------------------------------------------------------
// per-view constants (host computes them in double with the reference's own expressions)
// ------------------------------------------------------------------------------------------
struct ViewConst {
    float  cb, sb;        // cos / sin of beta (radians from degrees)           bp3d20.cpp:87-88
    float  ca, sa;        // 1/sqrt(1+tan^2(beta_deg)), tan(beta_deg)/sqrt(..)  bp3d20.cpp:137 (Q12)
    double cbd, sbd;      // cos(M_PI*beta/180), sin(M_PI*beta/180)
    double cnb, snb;      // cos(-1*M_PI*beta/180), sin(-1*M_PI*beta/180)       bp3d20.cpp:103-104
    double pxd, pyd;      // primary_x, primary_y                               bp3d20.cpp:87-88
};

static void make_view_consts(const monte_fdk_geom &g, std::vector<ViewConst> &out) {
    out.resize(g.n_views);
    for (int v = 0; v < g.n_views; v++) {
        double beta = g.angle0_deg + g.angle_step_deg * (double)v;
        float start_x = (float)(-g.dso), start_y = 0;
        ViewConst c;
        c.cbd = cos(M_PI * beta / 180);
        c.sbd = sin(M_PI * beta / 180);
        c.cnb = cos(-1 * M_PI * beta / 180);
        c.snb = sin(-1 * M_PI * beta / 180);
        c.pxd = start_x * c.cbd - start_y * c.sbd;
        c.pyd = start_x * c.sbd + start_y * c.cbd;
        double T = tan(beta);                       // degrees taken as radians, as shipped
        double inv = 1.0 / sqrt(1 + T * T);
        c.cb = (float)c.cbd; c.sb = (float)c.sbd;
        c.ca = (float)inv;   c.sa = (float)(T * inv);
        out[v] = c;
    }
}

static int check_geom(const monte_fdk_geom *g) {
    MONTE_ARG(g != nullptr, "fdk: geom is NULL");
    MONTE_ARG(g->n_views > 0 && g->nu > 0 && g->nv > 0, "fdk: n_views/nu/nv must be positive");
    MONTE_ARG(g->du > 0 && g->dv > 0 && g->vox > 0, "fdk: du/dv/vox must be positive");
    MONTE_ARG(g->nx > 0 && g->ny > 0 && g->nz > 0, "fdk: volume dims must be positive");
    MONTE_ARG(0 <= g->s_begin && g->s_begin <= g->s_end && g->s_end <= g->nx, "fdk: bad s ROI");
    MONTE_ARG(0 <= g->t_begin && g->t_begin <= g->t_end && g->t_end <= g->ny, "fdk: bad t ROI");
    MONTE_ARG(0 <= g->z_begin && g->z_begin <= g->z_end && g->z_end <= g->nz, "fdk: bad z ROI");
    MONTE_ARG(g->weight_mode == MONTE_FDK_REFERENCE || g->weight_mode == MONTE_FDK_TEXTBOOK,
              "fdk: unknown weight_mode %d", g->weight_mode);
    MONTE_ARG(g->dso > 0 && g->dsd > 0, "fdk: dso/dsd must be positive");
    MONTE_ARG((size_t)g->n_views * g->nv + 2 < (size_t)1 << 31, "fdk: too many projection rows");
    return MONTE_OK;
}

// ------------------------------------------------------------------------------------------
// weight + Ram-Lak filter
// ------------------------------------------------------------------------------------------
constexpr int FT_D = 8;          // axial rows (d) per CTA
constexpr int FT_NB = 16;        // outputs of one parity per task
constexpr int FT_THREADS = 256;

struct FilterParams {
    const float  *map;      // [views][nu][nv]
    const double *wtab;     // [nu][nv] cosine weights (double, host-computed)
    const float  *taps;     // odd-offset taps, index (n + noff) >> 1
    float        *out;      // padded rows
    int nu, nv, pitch;      // pitch of `out` rows in floats
    int view_begin;
    int nu_pad;             // zero-padded extent of a staged row (multiple of 16, >= nu+16)
    int in_pitch;           // smem row pitch, == 2 (mod 32)
    int noff, tap_len;
    float center;           // filter_scale * 0.25
};

__global__ void __launch_bounds__(FT_THREADS)
fdk_weight_filter_kernel(const FilterParams p) {
    MONTE_DYN_SMEM(float, smem);
    float *s_in  = smem;                                  // [FT_D][in_pitch]
    float *s_out = s_in + FT_D * p.in_pitch;              // [FT_D][in_pitch]
    float *s_tap = s_out + FT_D * p.in_pitch;             // [tap_len]
    const int tid = threadIdx.x;
    const int v = p.view_begin + blockIdx.y;
    const int d0 = blockIdx.x * FT_D;

    for (int m = tid; m < p.tap_len; m += FT_THREADS) s_tap[m] = p.taps[m];
    // step 1 fused into the load: map_w = float(double(map) * w)     (bp3d20.cpp:40)
    const float *mv = p.map + (size_t)v * p.nu * p.nv;
    for (int idx = tid; idx < FT_D * p.nu_pad; idx += FT_THREADS) {
        const int c = idx / FT_D, dl = idx % FT_D, d = d0 + dl;
        float val = 0.f;
        if (c < p.nu && d < p.nv) {
            const size_t i = (size_t)c * p.nv + d;
            val = (float)((double)__ldg(mv + i) * __ldg(p.wtab + i));
        }
        s_in[dl * p.in_pitch + c] = val;
    }
    __syncthreads();

    // step 2: out[b] = sum_c in[c]*tap(c-b); even offsets are zero, offset 0 is the centre tap.
    // A task owns FT_NB outputs of one parity b_j = b0 + 2j and walks the inputs of the other
    // parity in ascending c (the reference's order, bp3d20.cpp:67), 8 inputs per trip:
    // 8 + (8 + FT_NB - 1) shared-memory loads feed 8*FT_NB FMAs.
    constexpr int SW = 2 * FT_NB;                       // outputs covered by a strip (both parities)
    constexpr int NT = 8 + FT_NB - 1;                   // taps needed per trip
    const int n_strips = (p.nu + SW - 1) / SW;
    const int n_tasks = FT_D * 2 * n_strips;
    for (int task = tid; task < n_tasks; task += FT_THREADS) {
        const int dl = task & 7, q = (task >> 3) & 1, strip = task >> 4;
        const int b0 = strip * SW + q;
        const float *row = s_in + dl * p.in_pitch;
        float acc[FT_NB];
#pragma unroll
        for (int j = 0; j < FT_NB; j++) acc[j] = row[b0 + 2 * j] * p.center;
        for (int c = q ^ 1; c < p.nu; c += 16) {
            const int mbase = (c - b0 - 2 * (FT_NB - 1) + p.noff) >> 1;
            float tp[NT], x[8];
#pragma unroll
            for (int k = 0; k < NT; k++) tp[k] = s_tap[mbase + k];
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = row[c + 2 * u];
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int j = 0; j < FT_NB; j++) acc[j] = fmaf(x[u], tp[u - j + FT_NB - 1], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < FT_NB; j++) s_out[dl * p.in_pitch + b0 + 2 * j] = acc[j];
    }
    __syncthreads();
    for (int idx = tid; idx < FT_D * p.nu; idx += FT_THREADS) {
        const int dl = idx / p.nu, b = idx - dl * p.nu, d = d0 + dl;
        if (d < p.nv) p.out[((size_t)v * p.nv + d) * p.pitch + b] = s_out[dl * p.in_pitch + b];
    }
}

// ------------------------------------------------------------------------------------------
// weight + Ram-Lak filter by FFT (wide detectors).  Same sums as the kernel above, evaluated as a
// zero-padded circular convolution of length L >= 2 nu (fft_core.cuh): a CTA of L/8 threads takes
// FFT_PAIRS pairs of neighbouring axial columns of one view; a pair is one complex sequence
// (column d + i column d+1), so one forward and one inverse transform filter two rows.  Cost per row
// ~ 5 L log2 L flop instead of nu^2 (C3: 12x fewer), rounding error ~3e-7 of the row maximum.
// ------------------------------------------------------------------------------------------

struct FftFilterParams {
    const float  *map;      // [views][nu][nv]
    const double *wtab;     // [nu][nv]
    const cpx    *tw;       // [L] exp(-2 pi i m / L)
    const float  *spec;     // [L] spectrum of the scaled taps / L (real: the taps are even)
    float        *out;      // padded rows
    int nu, nv, pitch, view_begin;
};

constexpr int fft_tile_pitch(int L) { return L / 2 + 4; }          // cpx per staged pair; +4 spreads the pairs over the banks
constexpr int FFT_PAIRS = 4;                                        // 8 was measured slower (2 CTAs/SM instead of 3)
constexpr size_t fft_filter_smem(int L) { return ((size_t)2 * fft_padded_len(L) + FFT_PAIRS * fft_tile_pitch(L)) * sizeof(cpx); }

template <int L>
__global__ void __launch_bounds__(L / 8, 8192 / L)
fdk_weight_filter_fft_kernel(const FftFilterParams p) {
    MONTE_DYN_SMEM(cpx, s_fft);
    constexpr int T = FftPlan<L>::THREADS, PL = fft_padded_len(L), TP = fft_tile_pitch(L);
    cpx *bufA = s_fft, *bufB = s_fft + PL, *tile = s_fft + 2 * PL;
    const int j = threadIdx.x;
    const int v = p.view_begin + blockIdx.y;
    const int d0 = blockIdx.x * (2 * FFT_PAIRS);
    FftTwiddles<L> w;
    w.load(p.tw, j);
    // step 1 fused into the load: map_w = float(double(map) * w) (bp3d20.cpp:40).  The 2*FFT_PAIRS columns
    // of this CTA are 32 contiguous bytes of every detector row: fetched once, weighted, and parked as
    // complex pairs (column d + i column d+1).
    {
        const float *mv = p.map + (size_t)v * p.nu * p.nv;
        float *tf = reinterpret_cast<float *>(tile);
        for (int idx = j; idx < p.nu * (2 * FFT_PAIRS); idx += T) {
            const int n = idx / (2 * FFT_PAIRS), c = idx % (2 * FFT_PAIRS), d = d0 + c;
            float val = 0.f;
            if (d < p.nv) {
                const size_t i = (size_t)n * p.nv + d;
                val = (float)((double)__ldg(mv + i) * __ldg(p.wtab + i));
            }
            tf[((c >> 1) * TP + n) * 2 + (c & 1)] = val;
        }
    }
    __syncthreads();
    for (int q = 0; q < FFT_PAIRS; q++) {
        const int d = d0 + 2 * q;
        if (d >= p.nv) break;                                            // uniform
        const bool two = d + 1 < p.nv;
        const cpx *tq = tile + q * TP;
        auto in = [&](int n) { return n < p.nu ? tq[n] : cpx{0.f, 0.f}; };
        float *row0 = p.out + ((size_t)v * p.nv + d) * p.pitch;
        auto out = [&](int n, cpx val) {
            if (n < p.nu) {
                row0[n] = val.x;
                if (two) row0[p.pitch + n] = val.y;
            }
        };
#pragma unroll
        for (int phase = 0; phase < 8; phase++) {
            fft_filter_phase<L>(phase, j, bufA, bufB, w, p.spec, in, out);
            __syncthreads();
        }
    }
}

// element [r][nu] = [r+1][0]; columns nu+1.. and the two trailing rows are zero.
__global__ void fdk_pad_kernel(float *f, int rows, int nu, int pitch, int r0, int r1) {
    const int r = r0 + blockIdx.x * blockDim.x + threadIdx.x;      // rows [r0, r1) of the rows+2 in the buffer
    if (r >= r1) return;
    float *row = f + (size_t)r * pitch;
    if (r >= rows) {
        for (int c = 0; c < pitch; c++) row[c] = 0.f;
        return;
    }
    row[nu] = (r + 1 < rows) ? f[(size_t)(r + 1) * pitch] : 0.f;
    for (int c = nu + 1; c < pitch; c++) row[c] = 0.f;
}

// pairs[r][c] = { f[r][c], f[r+1][c] - f[r][c] }: the two rows a bilinear fetch needs sit in one aligned
// 8-byte element, so a voxel update issues 2 x LDG.64 instead of 4 x LDG.32 (the gathers were LSU-issue
// bound), and the row difference the interpolation starts with (same fp32 subtraction, same rounding as
// in the hot loop before) is taken once per texel here instead of once per voxel update there
// Only rows [b_lo, b_hi) of every view are paired: the band a z-slab projects onto (all rows for a
// whole-volume call, ~1/N of them for one of N multi-GPU slabs).
// fdk_pair_gather_kernel builds them, from the local padded rows or straight from the devices that filtered the views
// (multi-device reconstruction): there the
// exchange of filtered projections IS this kernel -- every device loads the detector-row band its z-slab reads out of
// its peers' memory (NVLink P2P, coalesced rows) while converting it to the pair layout, so no filtered view is ever
// copied, packed or stored twice.  src.base[o] is the (virtual) base of owner o's padded-row layout: row R of the
// global layout lives at base[o] + R * pitch for the views [v_end[o-1], v_end[o]) owner o filtered.  The pad rules of
// fdk_pad_kernel (element [r][nu] = [r+1][0], columns beyond and rows past the last view = 0) are applied on the fly.
// (A device filters several chunks of views, interleaved with the other devices' chunks so that the reconstruction can
// start while later chunks are still being uploaded: the layout is described per chunk = segment.)
constexpr int PAIR_MAX_SEG = MAX_DEV * 4;
struct PairSrc {
    const float *base[PAIR_MAX_SEG];   // segment o holds views [v_end[o-1], v_end[o]) (segments may be empty)
    int v_end[PAIR_MAX_SEG];
    int n;
};
__device__ __forceinline__ float pair_fetch(const PairSrc &src, int R, int c, int rows, int nv, int nu, int pitch, int seg0) {
    if (c > nu || R >= rows) return 0.f;
    if (c == nu) { R += 1; c = 0; if (R >= rows) return 0.f; }
    const int v = R / nv;
    int o = seg0;                                   // the segment of the launch's first view: the search starts there
    while (o < src.n - 1 && v >= src.v_end[o]) o++;
    return src.base[o][(size_t)R * pitch + c];
}
// A thread walks PAIR_RPT consecutive rows of one column: every texel is fetched once (+ one per strip) instead of
// twice (as row R and as the partner of row R - 1) -- the fetches are NVLink peer loads between devices.
// ONE launch converts everything a view chunk needs (one host thread feeds up to eight devices: launches count):
//   z <  n_v : view view_lo + z -- strip 0 = rows [0, a_hi) (what the previous view's last row reaches into; a_hi = 0:
//              the band starts at row 0 and covers them), strips 1.. = the band [b_lo, b_hi)
//   z == n_v : the view after the chunk -- rows [0, x_hi) only: 4 rows of the next view, or the two zero rows after the
//              last view of all
constexpr int PAIR_RPT = 8;
__global__ void fdk_pair_gather_kernel(const PairSrc src, float2 *__restrict__ pairs, int view_lo, int n_v, int nv, int nu,
                                       int a_hi, int b_lo, int b_hi, int x_hi, int rows, int pitch, int seg0) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= pitch) return;
    const bool extra = (int)blockIdx.z == n_v;
    int r_first, r_last;
    if (blockIdx.x == 0) { r_first = 0; r_last = extra ? x_hi : a_hi; }
    else {
        if (extra) return;
        r_first = b_lo + ((int)blockIdx.x - 1) * PAIR_RPT;
        r_last = min(r_first + PAIR_RPT, b_hi);
    }
    if (r_first >= r_last) return;
    int R = (view_lo + (int)blockIdx.z) * nv + r_first;
    float a = pair_fetch(src, R, c, rows, nv, nu, pitch, seg0);
    for (int r = r_first; r < r_last; r++, R++) {
        if (R >= rows + 2) return;
        const float b = pair_fetch(src, R + 1, c, rows, nv, nu, pitch, seg0);
        pairs[(size_t)R * pitch + c] = make_float2(a, b - a);
        a = b;
    }
}

__global__ void fdk_unpad_kernel(const float *f, float *dense, size_t rows, int nu, int pitch) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * nu) return;
    const size_t r = i / nu;
    const int c = (int)(i - r * nu);
    dense[i] = f[r * pitch + c];
}

// ------------------------------------------------------------------------------------------
// backprojection
// ------------------------------------------------------------------------------------------
constexpr int BP_TX = 32, BP_TY = 8;   // threads: 32 along s (x-fastest, coalesced), 8 along t
constexpr int BP_DEFAULT_VARIANT = 0;  // MONTE_BP_VARIANT overrides: 0 L1 gathers, 10 / 11 footprint staged in shared memory (4 / 3 CTAs per SM)

struct BpParams {
    const float2 *pairs;        // padded rows [n_views*nv + 2][pitch] as vertical pairs (fdk_pair_gather_kernel)
    const ViewConst *vc;
    float *vol;                 // slab base: slice z_lo
    int n_views, nu, nv, pitch;
    int accumulate;                   // start from the stored volume (view chunks after the first)
    int nx, ny;
    int s_begin, s_end, t_begin, t_end, z_lo, z_hi;   // z range of this launch (absolute)
    int roi_z_begin, roi_z_end;                        // geometry ROI in z (outside -> 0)
    float x0, y0, z0, vox;
    float dso, dsd, half_u, half_v, inv_du, inv_dv;
    float wd, wd2;              // weight_dist and its square (REFERENCE)
    float out_fac;              // beta_span*2*pi/360 * out_scale*out_scale2 (REFERENCE) or 0.5*dbeta
    float eps_u, eps_v;         // half-widths of the bands re-evaluated in double
    float kmin;                 // smallest magnification dsd / r_x any column of the ROI can have (0.99 x, for the z-block early-out)
    float eps_ts;               // |tmp_s| below this: its sign is re-evaluated in double
    int textbook, coord_mode;
    int mask_cs, mask_ct, mask_cz;
    long long mask_r2;
    // doubles for the exact edge path
    double x0d, y0d, z0d, voxd, dsdd, half_ud, half_vd, inv_dud, inv_dvd, nu_half, nv_half;
};

// The reference decides in double whether a voxel's projection is on the detector
// (bp3d20.cpp:116) and the decision is discontinuous, so fp32 values inside a narrow band
// around the edge are re-evaluated with the reference's exact double expressions (no FMA
// contraction).  Returns false if the reference skips this (voxel, view).
__device__ __noinline__ bool bp_exact(const BpParams &p, const ViewConst &c, int s, int t, int z,
                                      float &xf, float &yf) {
    double X = __dadd_rn(p.x0d, __dmul_rn((double)s, p.voxd));
    double Y = __dsub_rn(p.y0d, __dmul_rn((double)t, p.voxd));
    double Z = __dsub_rn(p.z0d, __dmul_rn((double)z, p.voxd));
    double pv0 = __dsub_rn(X, c.pxd), pv1 = __dsub_rn(Y, c.pyd);
    double r0 = __dsub_rn(__dmul_rn(pv0, c.cnb), __dmul_rn(pv1, c.snb));
    double r1 = __dadd_rn(__dmul_rn(pv0, c.snb), __dmul_rn(pv1, c.cnb));
    double to_det = __ddiv_rn(p.dsdd, r0);
    double u = __dmul_rn(r1, to_det), w = __dmul_rn(Z, to_det);
    if (fabs(u) > p.half_ud || fabs(w) > p.half_vd) return false;
    double y, x;
    if (p.coord_mode == MONTE_FDK_COORD_SCALE_BEFORE) {
        y = -__dsub_rn(__dmul_rn(u, p.inv_dud), p.nu_half);
        x = -__dsub_rn(__dmul_rn(w, p.inv_dvd), p.nv_half);
    } else {
        y = __dmul_rn(-p.inv_dud, __dsub_rn(u, p.half_ud));
        x = __dmul_rn(-p.inv_dvd, __dsub_rn(w, p.half_vd));
    }
    if (!(0 <= x && x <= p.nv && 0 <= y && y <= p.nu)) { xf = -1.f; yf = -1.f; return true; }
    xf = (float)x; yf = (float)y;
    return true;
}

// sign(tmp_s) flips the sign of d (bp3d20.cpp:140-142) and tmp_s crosses zero on a line through
// the rotation centre; the reference's double rounding decides there, so reproduce it.
__device__ __noinline__ bool bp_exact_ts_negative(const BpParams &p, const ViewConst &c, int s, int t) {
    double X = __dadd_rn(p.x0d, __dmul_rn((double)s, p.voxd));
    double Y = __dsub_rn(p.y0d, __dmul_rn((double)t, p.voxd));
    return __dsub_rn(__dmul_rn(X, c.cbd), __dmul_rn(Y, c.sbd)) < 0;
}

// the same four texels out of the pair layout: pairs[r][c].x IS f[r][c], so rows xi and xi + 1 give the bits
// bp_bilinear reads from the plain rows (the band a slab pairs always holds row xi + 1 of an on-detector xi)
__device__ __forceinline__ float bp_bilinear_pairs(const float2 *__restrict__ f, int pitch, float x, float y) {
    const int xi = (int)x, yi = (int)y;
    const float fx = x - (float)xi, fy = y - (float)yi;
    const float2 *q = f + (size_t)xi * pitch + yi;
    const float a = __ldg(&q->x), b = __ldg(&(q + 1)->x), c = __ldg(&(q + pitch)->x), d = __ldg(&(q + pitch + 1)->x);
    const float lo = fmaf(fx, c - a, a);      // (1-fx)*a + fx*c
    const float hi = fmaf(fx, d - b, b);
    return fmaf(fy, hi - lo, lo);
}

// exact contribution of one band voxel (rare path, kept out of line and out of the hot loop)
__device__ __noinline__ float bp_fix_one(const BpParams &p, const ViewConst &c, int s, int t, int z,
                                         const float2 *__restrict__ fpv, float wgt) {
    float xe, ye;
    if (!bp_exact(p, c, s, t, z, xe, ye) || xe < 0.f) return 0.f;
    xe = fminf(fmaxf(xe, 0.f), (float)p.nv);
    ye = fminf(fmaxf(ye, 0.f), (float)p.nu);
    return wgt * bp_bilinear_pairs(fpv, p.pitch, xe, ye);
}

// Hot loop structure per view: (1) everything that depends on (s,t,view) only; (2) for a batch of
// ZB z-slices: coordinates, then all 4*ZB gathers issued back to back (unconditional, clamped
// addresses, no branch or call in between so they overlap), then the bilinear FMAs with a select;
// (3) voxels inside the narrow bands around the detector edge contribute nothing in (2) and are
// added exactly afterwards.
template <int ZT, int ZB, int MINB>
__global__ void __launch_bounds__(BP_TX *BP_TY, MINB)
fdk_backproject_kernel(const __grid_constant__ BpParams p) {
    // a warp covers an 8 (s) x 4 (t) patch of columns, not 32 x 1: its detector footprint per
    // gather is ~2x narrower (fewer 128-byte lines per LDG), stores stay full 32-byte sectors
    const int tid = threadIdx.y * BP_TX + threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int s = p.s_begin + blockIdx.x * BP_TX + (wid & 3) * 8 + (lane & 7);
    const int t = p.t_begin + blockIdx.y * BP_TY + (wid >> 2) * 4 + (lane >> 3);
    // z blocks are aligned to absolute multiples of ZT so that the per-slice increments below round
    // identically however the volume is cut into slabs (multi-GPU slabs == single launch, bit for bit)
    const int zb = (p.z_lo / ZT) * ZT + blockIdx.z * ZT;
    if (s >= p.s_end || t >= p.t_end) return;
    {
        // z-block entirely above or below the detector even at the smallest magnification any column can
        // have: no view contributes (bp3d20.cpp:116), the stored zeros / partial sums stay as they are
        const float Za = fmaf(-p.vox, (float)zb, p.z0), Zb = fmaf(-p.vox, (float)(zb + ZT - 1), p.z0);
        if (Za * Zb > 0.f && fminf(fabsf(Za), fabsf(Zb)) * p.kmin > p.half_v + 2.f * p.eps_v) return;
    }

    const float X = fmaf(p.vox, (float)s, p.x0);
    const float Y = fmaf(-p.vox, (float)t, p.y0);
    const float Z0 = fmaf(-p.vox, (float)zb, p.z0);
    float acc[ZT];
#pragma unroll
    for (int i = 0; i < ZT; i++) acc[i] = 0.f;
    if (p.accumulate) {                              // later view chunks continue the stored fp32 partial sums
#pragma unroll
        for (int i = 0; i < ZT; i++) {
            const int z = zb + i;
            if (z >= p.z_lo && z < p.z_hi) acc[i] = p.vol[((size_t)(z - p.z_lo) * p.ny + t) * p.nx + s];
        }
    }
    const float xoff = p.half_v * p.inv_dv;
    const float hv_in = p.half_v - p.eps_v, hu_in = p.half_u - p.eps_u;
    const float nvf = (float)p.nv, nuf = (float)p.nu;

    for (int v = 0; v < p.n_views; v++) {
        const float4 cv = __ldg(reinterpret_cast<const float4 *>(p.vc + v));     // cb, sb, ca, sa
        const float cb = cv.x, sb = cv.y;
        const float rx = fmaf(X, cb, fmaf(Y, sb, p.dso));       // distance from the source along the axis
        const float ry = fmaf(Y, cb, -X * sb);
        const float k = __fdividef(p.dsd, rx);
        const float u = k * ry;
        const float au = fabsf(u);
        if (au > p.half_u + p.eps_u) continue;                  // bp3d20.cpp:116
        const bool u_ok = au <= hu_in;                          // else: inside the u band -> exact path for the column
        float wgt;
        if (!p.textbook) {
            const float ts = fmaf(X, cb, -Y * sb);               // bp3d20.cpp:134-142
            const float tt = fmaf(X, sb, Y * cb);
            float d = fabsf(fmaf(ts, cv.z, -tt * cv.w));
            bool neg = ts < 0.f;
            if (fabsf(ts) < p.eps_ts) neg = bp_exact_ts_negative(p, p.vc[v], s, t);
            d = neg ? -d : d;
            const float e = p.wd - d;
            wgt = __fdividef(p.wd2, e * e) * p.out_fac;
        } else {
            wgt = __fdividef(p.dso * p.dso, rx * rx) * p.out_fac;
        }
        const float y = fminf(fmaxf((p.half_u - u) * p.inv_du, 0.f), nuf);
        const int yi = (int)y;
        const float fy = y - (float)yi;
        // warp-uniform row pointers + one 32-bit element offset per voxel: one IMAD and two
        // IMAD.WIDE per update instead of 64-bit add/shift chains on the ALU pipe
        const float2 *__restrict__ fp = p.pairs + (size_t)v * p.nv * p.pitch;
        // column yi of this view's rows as one opaque 64-bit value: otherwise the compiler keeps the
        // warp-uniform view base apart and re-adds it for every slice (IADD3 + IADD3.X per update)
        unsigned long long fcol = reinterpret_cast<unsigned long long>(fp + yi);
        asm volatile("" : "+l"(fcol));
        const unsigned pitch_bytes = (unsigned)p.pitch * (unsigned)sizeof(float2);
        const float kz = k * p.inv_dv;
        const float kzv = kz * p.vox, kv = k * p.vox;            // per-slice increments of x and w
        const float x0v = fmaf(-kz, Z0, xoff), w0v = k * Z0;
        unsigned band = 0;
        // w is linear in the slice index: if both ends of the column are safely inside the detector
        // every slice is, and the per-slice edge tests and clamps can be dropped altogether
        const float w_last = fmaf(-kv, (float)(ZT - 1), w0v);
        // whole column above or below the detector for this view (bp3d20.cpp:116 skips every slice)
        if (w0v * w_last > 0.f && fminf(fabsf(w0v), fabsf(w_last)) > p.half_v + p.eps_v) continue;
        if (u_ok && fmaxf(fabsf(w0v), fabsf(w_last)) <= hv_in) {
#pragma unroll
            for (int h = 0; h < ZT; h += ZB) {
                float fx[ZB], va[ZB], vb[ZB], vc2[ZB], vd[ZB];
#pragma unroll
                for (int i = 0; i < ZB; i++) {
                    // x >= -ulp by construction; truncation maps that to row 0 with fx = -ulp
                    const float x = fmaf(kzv, (float)(h + i), x0v);
                    const int xi = (int)x;
                    fx[i] = x - (float)xi;
                    // one IMAD.WIDE: 64-bit (warp-uniform row base + column) + xi * row pitch in bytes
                    const float2 *__restrict__ q = reinterpret_cast<const float2 *>(fcol + (unsigned long long)(unsigned)xi * pitch_bytes);
                    const float2 p0 = __ldg(q), p1 = __ldg(q + 1);
                    va[i] = p0.x; vc2[i] = p0.y; vb[i] = p1.x; vd[i] = p1.y;
                }
#pragma unroll
                for (int i = 0; i < ZB; i++) {
                    const float lo = fmaf(fx[i], vc2[i], va[i]);              // a + fx*(c-a)
                    const float hi = fmaf(fx[i], vd[i], vb[i]);
                    acc[h + i] = fmaf(wgt, fmaf(fy, hi - lo, lo), acc[h + i]);
                }
            }
            continue;
        }
#pragma unroll
        for (int h = 0; h < ZT; h += ZB) {
            float fx[ZB], va[ZB], vb[ZB], vc2[ZB], vd[ZB];
            bool ok[ZB];
#pragma unroll
            for (int i = 0; i < ZB; i++) {
                const float fi = (float)(h + i);
                const float w = fmaf(-kv, fi, w0v);             // k*(Z0 - vox*i)
                const float aw = fabsf(w);
                float x = fmaf(kzv, fi, x0v);
                ok[i] = u_ok && aw <= hv_in;
                if (fabsf(aw - p.half_v) < p.eps_v || (!u_ok && aw < p.half_v + p.eps_v)) band |= 1u << (h + i);
                x = fminf(fmaxf(x, 0.f), nvf);
                const int xi = (int)x;
                fx[i] = x - (float)xi;
                const float2 *__restrict__ q = reinterpret_cast<const float2 *>(fcol + (unsigned long long)(unsigned)xi * pitch_bytes);
                const float2 p0 = __ldg(q), p1 = __ldg(q + 1);
                va[i] = p0.x; vc2[i] = p0.y; vb[i] = p1.x; vd[i] = p1.y;
            }
#pragma unroll
            for (int i = 0; i < ZB; i++) {
                const float lo = fmaf(fx[i], vc2[i], va[i]);
                const float hi = fmaf(fx[i], vd[i], vb[i]);
                const float val = fmaf(fy, hi - lo, lo);
                acc[h + i] = fmaf(ok[i] ? wgt : 0.f, val, acc[h + i]);
            }
        }
        if (band) {                                             // rare: exact double evaluation
#pragma unroll
            for (int i = 0; i < ZT; i++)
                if (band & (1u << i)) acc[i] += bp_fix_one(p, p.vc[v], s, t, zb + i, p.pairs + (size_t)v * p.nv * p.pitch, wgt);
        }
    }
#pragma unroll
    for (int i = 0; i < ZT; i++) {
        const int z = zb + i;
        if (z >= p.z_lo && z < p.z_hi) {
            float r = acc[i];
            if (z < p.roi_z_begin || z >= p.roi_z_end) r = 0.f;
            if (p.mask_r2 >= 0) {
                const long long dz = z - p.mask_cz, dt = t - p.mask_ct, ds = s - p.mask_cs;
                if (dz * dz + dt * dt + ds * ds > p.mask_r2) r = 0.f;
            }
            p.vol[((size_t)(z - p.z_lo) * p.ny + t) * p.nx + s] = r;
        }
    }
}

#ifndef MONTE_EMU
// ------------------------------------------------------------------------------------------
// Backprojection with the detector footprint of a CTA staged in shared memory by TMA.
// A CTA owns a brick of 16 x 16 columns x ZT slices.  Per view its voxels project into a small rectangle of the
// row-pair layout (C3: ~50 columns x ~42 rows = 17 KB) that the 4096 updates of the brick re-read ~5 times; the
// kernel above takes those re-reads from L1 (LDG.64 x 2 per update: the L1 tag/data path was the busiest unit,
// 83 %).  Here ONE thread issues ONE `cp.async.bulk.tensor.2d` per view (a 2-D tensor map over the pair rows; the
// hardware clips the box at the buffer's edges), three stages deep, completion counted on `full` mbarriers; a stage
// is handed back through an `empty` mbarrier on which every warp arrives when it has finished the view, so there
// is no CTA-wide barrier in the loop and warps may drift a view apart.  The update reads its two texels with
// LDS.64.  Same arithmetic on the same texels: the volume is bit-identical.  Columns whose footprint is not safely
// inside the detector, or falls outside the staged box, take the global-memory path of the kernel above, view by
// view.  (v1 of this kernel issued one non-tensor bulk copy per tile row from the lanes of warp 0: the compiler
// serialises such copies lane by lane -- UBLKCP takes uniform registers -- and with a __syncthreads per view
// the whole CTA waited for that warp: 85 ms against 56 ms; profiles/r02_fdk_smem_v1_ncu_full.txt.)
// ------------------------------------------------------------------------------------------
constexpr int BPS_T = 16;                        // columns per CTA along s and along t
constexpr int BPS_STAGES = 3;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <int ZT, int ZB, int MINB>
__global__ void __launch_bounds__(BPS_T *BPS_T, MINB)
fdk_backproject_smem_kernel(const __grid_constant__ BpParams p, const __grid_constant__ CUtensorMap tmap, const int C, const int R,
                            const int stage_elems) {
    extern __shared__ __align__(128) unsigned char bps_smem[];
    float2 *tile = reinterpret_cast<float2 *>(bps_smem);                        // [BPS_STAGES][stage_elems >= R * C] row pairs
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(tile + BPS_STAGES * stage_elems);   // full[3], empty[3]
    int2 *s_org = reinterpret_cast<int2 *>(bars + 2 * BPS_STAGES);              // [3] {c0, r0}
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // a warp covers an 8 (s) x 4 (t) patch of columns, the CTA 2 x 4 such patches
    const int s_cta = p.s_begin + blockIdx.x * BPS_T, t_cta = p.t_begin + blockIdx.y * BPS_T;
    const int s = s_cta + (wid & 1) * 8 + (lane & 7);
    const int t = t_cta + (wid >> 1) * 4 + (lane >> 3);
    const int zb = (p.z_lo / ZT) * ZT + blockIdx.z * ZT;
    const bool valid = s < p.s_end && t < p.t_end;
    {
        const float Za = fmaf(-p.vox, (float)zb, p.z0), Zb = fmaf(-p.vox, (float)(zb + ZT - 1), p.z0);
        if (Za * Zb > 0.f && fminf(fabsf(Za), fabsf(Zb)) * p.kmin > p.half_v + 2.f * p.eps_v) return;      // CTA-uniform
    }
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8u * BPS_STAGES, tile0 = smem_u32(tile);
    const uint32_t stage_bytes = (uint32_t)stage_elems * (uint32_t)sizeof(float2);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < BPS_STAGES; i++) { mbar_init(bar_full + 8u * i, 1); mbar_init(bar_empty + 8u * i, BPS_T * BPS_T / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const float X = fmaf(p.vox, (float)s, p.x0);
    const float Y = fmaf(-p.vox, (float)t, p.y0);
    const float Z0 = fmaf(-p.vox, (float)zb, p.z0);
    const float xoff = p.half_v * p.inv_dv;
    // stage the footprint of view v (one thread): the box origin from the brick's corners -- the extremes of the
    // magnification and of the transaxial coordinate over the brick are taken there
    auto issue = [&](int v) {
        const int b = v % BPS_STAGES;
        const float Xa = fmaf(p.vox, (float)s_cta, p.x0), Xb = fmaf(p.vox, (float)min(s_cta + BPS_T - 1, p.s_end - 1), p.x0);
        const float Ya = fmaf(-p.vox, (float)t_cta, p.y0), Yb = fmaf(-p.vox, (float)min(t_cta + BPS_T - 1, p.t_end - 1), p.y0);
        const float Zl = fmaf(-p.vox, (float)(ZT - 1), Z0);
        const float4 cv = __ldg(reinterpret_cast<const float4 *>(p.vc + v));
        float kmn = 1e30f, kmx = -1e30f, ymn = 1e30f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float Xc = (q & 1) ? Xb : Xa, Yc = (q & 2) ? Yb : Ya;
            const float rx = fmaf(Xc, cv.x, fmaf(Yc, cv.y, p.dso)), ry = fmaf(Yc, cv.x, -Xc * cv.y);
            const float k = __fdividef(p.dsd, rx);
            kmn = fminf(kmn, k); kmx = fmaxf(kmx, k); ymn = fminf(ymn, (p.half_u - k * ry) * p.inv_du);
        }
        const float xa = fmaf(-kmn * p.inv_dv, Z0, xoff), xb = fmaf(-kmx * p.inv_dv, Z0, xoff);
        const float xc = fmaf(-kmn * p.inv_dv, Zl, xoff), xd = fmaf(-kmx * p.inv_dv, Zl, xoff);
        // one texel of slack against fp32 rounding differences with the per-thread coordinates; columns from an even
        // index (16-byte aligned box rows).  Threads whose texels fall outside the box use the global path.
        const int c0 = ((int)floorf(fmaxf(ymn, 0.f)) - 1) & ~1;
        const int r0 = (int)floorf(fmaxf(fminf(fminf(xa, xb), fminf(xc, xd)), 0.f)) - 1;
        s_org[b] = make_int2(c0, r0);
        mbar_arrive_expect_tx(bar_full + 8u * b, (uint32_t)(R * C) * (uint32_t)sizeof(float2));
        tma_load_2d(tile0 + (uint32_t)b * stage_bytes, &tmap, 2 * c0, v * p.nv + r0, bar_full + 8u * b);
    };
    if (tid == 0)
        for (int v = 0; v < BPS_STAGES && v < p.n_views; v++) issue(v);

    float acc[ZT];
#pragma unroll
    for (int i = 0; i < ZT; i++) acc[i] = 0.f;
    if (p.accumulate && valid) {
#pragma unroll
        for (int i = 0; i < ZT; i++) {
            const int z = zb + i;
            if (z >= p.z_lo && z < p.z_hi) acc[i] = p.vol[((size_t)(z - p.z_lo) * p.ny + t) * p.nx + s];
        }
    }
    const float hv_in = p.half_v - p.eps_v, hu_in = p.half_u - p.eps_u;
    const float nvf = (float)p.nv, nuf = (float)p.nu;

    for (int v = 0; v < p.n_views; v++) {
        const int b = v % BPS_STAGES;
        // the warps take turns as producer: at the start of view v, once every warp has handed back the stage of view
        // v - 1, that stage is refilled with view v + 2 (two views of lead time)
        if (v >= 1 && v + BPS_STAGES - 1 < p.n_views && wid == (v & 7) && lane == 0) {
            const int u = v - 1, ub = u % BPS_STAGES;
            const uint32_t par = (uint32_t)(u / BPS_STAGES) & 1u;
            mbar_wait(bar_empty + 8u * ub, par);
            mbar_wait(bar_full + 8u * ub, par);          // (its copy has landed even if no thread read it)
            issue(v + BPS_STAGES - 1);
        }
        // EVERY thread passes through the full barrier of every view, also those that will not read the box: it bounds
        // the run-ahead of a warp to the staged views (a stage is re-armed only after all warps handed it back), which
        // keeps the arrivals on `empty` in their own phase and makes the parity waits unambiguous
        mbar_wait(bar_full + 8u * b, (uint32_t)(v / BPS_STAGES) & 1u);
        const int2 org = s_org[b];
        do {
            if (!valid) break;
            const float4 cv = __ldg(reinterpret_cast<const float4 *>(p.vc + v));     // cb, sb, ca, sa
            const float cb = cv.x, sb = cv.y;
            const float rx = fmaf(X, cb, fmaf(Y, sb, p.dso));
            const float ry = fmaf(Y, cb, -X * sb);
            const float k = __fdividef(p.dsd, rx);
            const float u = k * ry;
            const float au = fabsf(u);
            if (au > p.half_u + p.eps_u) break;                      // bp3d20.cpp:116
            const bool u_ok = au <= hu_in;
            float wgt;
            if (!p.textbook) {
                const float ts = fmaf(X, cb, -Y * sb);               // bp3d20.cpp:134-142
                const float tt = fmaf(X, sb, Y * cb);
                float d = fabsf(fmaf(ts, cv.z, -tt * cv.w));
                bool neg = ts < 0.f;
                if (fabsf(ts) < p.eps_ts) neg = bp_exact_ts_negative(p, p.vc[v], s, t);
                d = neg ? -d : d;
                const float e = p.wd - d;
                wgt = __fdividef(p.wd2, e * e) * p.out_fac;
            } else {
                wgt = __fdividef(p.dso * p.dso, rx * rx) * p.out_fac;
            }
            const float y = fminf(fmaxf((p.half_u - u) * p.inv_du, 0.f), nuf);
            const int yi = (int)y;
            const float fy = y - (float)yi;
            const float2 *__restrict__ fp = p.pairs + (size_t)v * p.nv * p.pitch;
            const float kz = k * p.inv_dv;
            const float kzv = kz * p.vox, kv = k * p.vox;
            const float x0v = fmaf(-kz, Z0, xoff), w0v = k * Z0;
            unsigned band = 0;
            const float w_last = fmaf(-kv, (float)(ZT - 1), w0v);
            if (w0v * w_last > 0.f && fminf(fabsf(w0v), fabsf(w_last)) > p.half_v + p.eps_v) break;
            if (u_ok && fmaxf(fabsf(w0v), fabsf(w_last)) <= hv_in) {
                const float x_last = fmaf(kzv, (float)(ZT - 1), x0v);
                const int xlo = (int)fminf(x0v, x_last) - org.y, xhi = (int)fmaxf(x0v, x_last) - org.y, yc = yi - org.x;
                if (yc >= 0 && yc + 1 < C && xlo >= 0 && xhi < R) {
                    const uint32_t tb = tile0 + (uint32_t)b * stage_bytes + (uint32_t)(yc - org.y * C) * (uint32_t)sizeof(float2);
#pragma unroll
                    for (int h = 0; h < ZT; h += ZB) {
                        float fx[ZB]; float2 p0[ZB], p1[ZB];
#pragma unroll
                        for (int i = 0; i < ZB; i++) {
                            const float x = fmaf(kzv, (float)(h + i), x0v);
                            const int xi = (int)x;
                            fx[i] = x - (float)xi;
                            const uint32_t a = tb + (uint32_t)(xi * C) * (uint32_t)sizeof(float2);
                            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(p0[i].x), "=f"(p0[i].y) : "r"(a));
                            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(p1[i].x), "=f"(p1[i].y) : "r"(a));
                        }
#pragma unroll
                        for (int i = 0; i < ZB; i++) {
                            const float lo = fmaf(fx[i], p0[i].y, p0[i].x);
                            const float hi = fmaf(fx[i], p1[i].y, p1[i].x);
                            acc[h + i] = fmaf(wgt, fmaf(fy, hi - lo, lo), acc[h + i]);
                        }
                    }
                    break;
                }
                // texels outside the staged box: the same updates from global memory
                unsigned long long fcol = reinterpret_cast<unsigned long long>(fp + yi);
                const unsigned pitch_bytes = (unsigned)p.pitch * (unsigned)sizeof(float2);
#pragma unroll
                for (int h = 0; h < ZT; h += ZB) {
                    float fx[ZB], va[ZB], vb[ZB], vc2[ZB], vd[ZB];
#pragma unroll
                    for (int i = 0; i < ZB; i++) {
                        const float x = fmaf(kzv, (float)(h + i), x0v);
                        const int xi = (int)x;
                        fx[i] = x - (float)xi;
                        const float2 *__restrict__ q = reinterpret_cast<const float2 *>(fcol + (unsigned long long)(unsigned)xi * pitch_bytes);
                        const float2 q0 = __ldg(q), q1 = __ldg(q + 1);
                        va[i] = q0.x; vc2[i] = q0.y; vb[i] = q1.x; vd[i] = q1.y;
                    }
#pragma unroll
                    for (int i = 0; i < ZB; i++) {
                        const float lo = fmaf(fx[i], vc2[i], va[i]);
                        const float hi = fmaf(fx[i], vd[i], vb[i]);
                        acc[h + i] = fmaf(wgt, fmaf(fy, hi - lo, lo), acc[h + i]);
                    }
                }
                break;
            }
            {
                unsigned long long fcol = reinterpret_cast<unsigned long long>(fp + yi);
                const unsigned pitch_bytes = (unsigned)p.pitch * (unsigned)sizeof(float2);
#pragma unroll
                for (int h = 0; h < ZT; h += ZB) {
                    float fx[ZB], va[ZB], vb[ZB], vc2[ZB], vd[ZB];
                    bool ok[ZB];
#pragma unroll
                    for (int i = 0; i < ZB; i++) {
                        const float fi = (float)(h + i);
                        const float w = fmaf(-kv, fi, w0v);
                        const float aw = fabsf(w);
                        float x = fmaf(kzv, fi, x0v);
                        ok[i] = u_ok && aw <= hv_in;
                        if (fabsf(aw - p.half_v) < p.eps_v || (!u_ok && aw < p.half_v + p.eps_v)) band |= 1u << (h + i);
                        x = fminf(fmaxf(x, 0.f), nvf);
                        const int xi = (int)x;
                        fx[i] = x - (float)xi;
                        const float2 *__restrict__ q = reinterpret_cast<const float2 *>(fcol + (unsigned long long)(unsigned)xi * pitch_bytes);
                        const float2 q0 = __ldg(q), q1 = __ldg(q + 1);
                        va[i] = q0.x; vc2[i] = q0.y; vb[i] = q1.x; vd[i] = q1.y;
                    }
#pragma unroll
                    for (int i = 0; i < ZB; i++) {
                        const float lo = fmaf(fx[i], vc2[i], va[i]);
                        const float hi = fmaf(fx[i], vd[i], vb[i]);
                        const float val = fmaf(fy, hi - lo, lo);
                        acc[h + i] = fmaf(ok[i] ? wgt : 0.f, val, acc[h + i]);
                    }
                }
                if (band) {
#pragma unroll
                    for (int i = 0; i < ZT; i++)
                        if (band & (1u << i)) acc[i] += bp_fix_one(p, p.vc[v], s, t, zb + i, p.pairs + (size_t)v * p.nv * p.pitch, wgt);
                }
            }
        } while (0);
        // this warp is done with the stage of view v
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8u * b);
    }
    // the last staged boxes may still be in flight: a CTA must not exit under an outstanding bulk copy
    if (tid == 0)
        for (int v = max(p.n_views - BPS_STAGES, 0); v < p.n_views; v++) mbar_wait(bar_full + 8u * (v % BPS_STAGES), (uint32_t)(v / BPS_STAGES) & 1u);
    if (!valid) return;
#pragma unroll
    for (int i = 0; i < ZT; i++) {
        const int z = zb + i;
        if (z >= p.z_lo && z < p.z_hi) {
            float r = acc[i];
            if (z < p.roi_z_begin || z >= p.roi_z_end) r = 0.f;
            if (p.mask_r2 >= 0) {
                const long long dz = z - p.mask_cz, dt = t - p.mask_ct, ds = s - p.mask_cs;
                if (dz * dz + dt * dt + ds * ds > p.mask_r2) r = 0.f;
            }
            p.vol[((size_t)(z - p.z_lo) * p.ny + t) * p.nx + s] = r;
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libmonte_gpu does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_pairs_tensor_map(CUtensorMap *out, const float2 *pairs, size_t rows, int pitch, int C, int R) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        MONTE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr));
        if (!sym || qr != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return MONTE_E_CUDA; }
        fn = (EncodeTiledFn)sym;
    }
    // the pair rows as a 2-D fp32 tensor [rows][2 * pitch]; box = 2C floats x R rows; reads outside it give zeros
    const cuuint64_t dims[2] = {(cuuint64_t)2 * pitch, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float2)};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * C), (cuuint32_t)R}, estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)pairs, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for %zu rows of pitch %d, box %d x %d", (int)r, rows, pitch, C, R); return MONTE_E_CUDA; }
    return MONTE_OK;
}
#endif   // !MONTE_EMU

// vol_zy[s][t][z] = vol_xy[z][t][s]  (32x32 tiles of the (z,s) plane for every t)
__global__ void fdk_transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int nx, int ny, int nz) {
    __shared__ float tile[32][33];
    const int t = blockIdx.z;
    int s = blockIdx.x * 32 + threadIdx.x, z = blockIdx.y * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8)
        if (s < nx && z + j < nz) tile[threadIdx.y + j][threadIdx.x] = in[((size_t)(z + j) * ny + t) * nx + s];
    __syncthreads();
    z = blockIdx.y * 32 + threadIdx.x;
    s = blockIdx.x * 32 + threadIdx.y;
    for (int j = 0; j < 32; j += 8)
        if (z < nz && s + j < nx) out[((size_t)(s + j) * ny + t) * nz + z] = tile[threadIdx.x][threadIdx.y + j];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct FdkCache {           // per-geometry device constants, rebuilt only when the geometry changes
    monte_fdk_geom g;
    bool valid = false;
    ViewConst *d_vc = nullptr;
    double *d_wtab = nullptr;
    float *d_taps = nullptr;
    cpx *d_tw = nullptr;        // FFT filter: twiddles and tap spectrum for length fft_len (0: direct only)
    float *d_spec = nullptr;
    int fft_len = 0;
    int noff = 0, tap_len = 0, nu_pad = 0, in_pitch = 0;
    size_t vc_cap = 0, wtab_cap = 0, taps_cap = 0, fft_cap = 0;
};
// one cache (and one set of events / raised function attributes) per bound device
static PerDev<FdkCache> g_fdk_pd;
#define g_fdk (g_fdk_pd.get())
// per-context state that must not survive monte_gpu_shutdown (a later monte_gpu_init may bind another device):
// function attributes already raised, and the events of the host-buffer pipeline
constexpr int FDK_MAXC = 16;
struct FdkDevState {
    bool fft_attr_set = false;
    size_t filter_smem_set = 0, bps_smem_set[2] = {0, 0};
    cudaEvent_t ev_up[FDK_MAXC] = {nullptr}, ev_slab[FDK_MAXC] = {nullptr}, ev_t[6] = {nullptr};
    // z-block streams of the backprojector (thin multi-GPU slabs, see backproject_views)
    cudaStream_t zs[8] = {nullptr};
    cudaEvent_t ev_zfork = nullptr, ev_zjoin[8] = {nullptr};
    // pair-conversion stream of the backprojector: the conversion (and, between devices, the NVLink gather) of view chunk
    // c + 1 runs underneath the backprojection of chunk c
    cudaStream_t ps = nullptr;
    cudaEvent_t ev_pfork = nullptr, ev_pair[32] = {nullptr};
};
static PerDev<FdkDevState> g_fdk_state;
#define g_fft_attr_set (g_fdk_state.get().fft_attr_set)
#define g_filter_smem_set (g_fdk_state.get().filter_smem_set)
#define g_ev_up (g_fdk_state.get().ev_up)
#define g_ev_slab (g_fdk_state.get().ev_slab)
#define g_ev_t (g_fdk_state.get().ev_t)
static void fdk_cleanup() {                       // called once per bound device, with that device current
    cudaFree(g_fdk.d_vc); cudaFree(g_fdk.d_wtab); cudaFree(g_fdk.d_taps); cudaFree(g_fdk.d_tw); cudaFree(g_fdk.d_spec);
    g_fdk = FdkCache();
    for (int i = 0; i < FDK_MAXC; i++) {
        if (g_ev_up[i]) cudaEventDestroy(g_ev_up[i]);
        if (g_ev_slab[i]) cudaEventDestroy(g_ev_slab[i]);
    }
    for (int i = 0; i < 6; i++) if (g_ev_t[i]) cudaEventDestroy(g_ev_t[i]);
    {
        FdkDevState &ds = g_fdk_state.get();
        for (int i = 0; i < 8; i++) { if (ds.zs[i]) cudaStreamDestroy(ds.zs[i]); if (ds.ev_zjoin[i]) cudaEventDestroy(ds.ev_zjoin[i]); }
        if (ds.ev_zfork) cudaEventDestroy(ds.ev_zfork);
        if (ds.ps) cudaStreamDestroy(ds.ps);
        if (ds.ev_pfork) cudaEventDestroy(ds.ev_pfork);
        for (int i = 0; i < 32; i++) if (ds.ev_pair[i]) cudaEventDestroy(ds.ev_pair[i]);
    }
    g_fdk_state.get() = FdkDevState();
}

// the geometry as a cache key: copied field by field into a zeroed struct, so that the padding bytes of a
// caller's hand-filled monte_fdk_geom (after nv, after nz, before mask_r2) cannot make equal geometries differ
static monte_fdk_geom canon_geom(const monte_fdk_geom *g) {
    monte_fdk_geom c;
    memset(&c, 0, sizeof(c));
    c.n_views = g->n_views; c.nu = g->nu; c.nv = g->nv; c.du = g->du; c.dv = g->dv;
    c.half_u = g->half_u; c.half_v = g->half_v; c.dso = g->dso; c.dsd = g->dsd;
    c.weight_dist = g->weight_dist; c.filter_scale = g->filter_scale; c.out_scale = g->out_scale; c.out_scale2 = g->out_scale2;
    c.angle0_deg = g->angle0_deg; c.angle_step_deg = g->angle_step_deg;
    c.nx = g->nx; c.ny = g->ny; c.nz = g->nz; c.vox = g->vox; c.x0 = g->x0; c.y0 = g->y0; c.z0 = g->z0;
    c.s_begin = g->s_begin; c.s_end = g->s_end; c.t_begin = g->t_begin; c.t_end = g->t_end; c.z_begin = g->z_begin; c.z_end = g->z_end;
    c.mask_cs = g->mask_cs; c.mask_ct = g->mask_ct; c.mask_cz = g->mask_cz; c.mask_r2 = g->mask_r2;
    c.weight_mode = g->weight_mode; c.coord_mode = g->coord_mode;
    return c;
}

static int fdk_prepare(const monte_fdk_geom *g, cudaStream_t st) {
    const monte_fdk_geom key = canon_geom(g);
    if (g_fdk.valid && memcmp(&g_fdk.g, &key, sizeof(key)) == 0) return MONTE_OK;
    at_shutdown(fdk_cleanup);
    g_fdk.valid = false;
    std::vector<ViewConst> vc;
    make_view_consts(*g, vc);
    // cosine weights, bp3d20.cpp:40 (REFERENCE: weight_dist=60; TEXTBOOK: Dsd)
    const bool textbook = g->weight_mode == MONTE_FDK_TEXTBOOK;
    const double wd = textbook ? g->dsd : g->weight_dist;
    std::vector<double> wtab((size_t)g->nu * g->nv);
    for (int zeta = 0; zeta < g->nu; zeta++)
        for (int pp = 0; pp < g->nv; pp++) {
            double a = -1 * (zeta * g->du) + g->half_u, b = (pp * g->dv) - g->half_v;
            wtab[(size_t)zeta * g->nv + pp] = wd / sqrt(pow(wd, 2) + pow(a, 2) + pow(b, 2));
        }
    // Ram-Lak taps for odd offsets, bp3d20.cpp:57 (float) times the filter scale (double product)
    const int nu_pad = ((g->nu + 2 * FT_NB - 1) / (2 * FT_NB)) * (2 * FT_NB) + 16;   // strips of 2*FT_NB outputs + 16 inputs of slack
    int noff = nu_pad + 2 * FT_NB + 1; if ((noff & 1) == 0) noff++;
    const int tap_len = (nu_pad + 2 + noff) / 2 + FT_NB + 8;
    const double scale = textbook ? g->dsd / (g->dso * g->du) : g->filter_scale;
    std::vector<float> taps(tap_len, 0.f);
    for (int m = 0; m < tap_len; m++) {
        int n = 2 * m - noff, an = n < 0 ? -n : n;
        if ((an & 1) && an <= g->nu - 1) {
            float ramp = (float)(-1. / pow(an * M_PI, 2));
            taps[m] = (float)(scale * (double)ramp);
        }
    }
    int in_pitch = nu_pad;
    while ((in_pitch & 31) != 2) in_pitch++;
    // FFT filter tables: circular taps g[0] = centre, g[+-n] = odd taps (the same float values the direct
    // kernel multiplies with), spectrum H[k] = sum_n g[n] cos(2 pi k n / L) / L in double
    const int fft_len = g->nu <= 512 ? 1024 : g->nu <= 1024 ? 2048 : g->nu <= 2048 ? 4096 : 0;
    std::vector<cpx> tw;
    std::vector<float> spec;
    if (fft_len) {
        const int L = fft_len;
        std::vector<double> ctab(L), gcirc(L, 0.0);
        tw.resize(L); spec.resize(L);
        for (int m = 0; m < L; m++) {
            ctab[m] = cos(2.0 * M_PI * (double)m / L);
            tw[m] = cpx{(float)ctab[m], (float)(-sin(2.0 * M_PI * (double)m / L))};
        }
        gcirc[0] = (double)(float)(scale * 0.25);
        for (int n = 1; n <= g->nu - 1; n += 2) {
            const float ramp = (float)(-1. / pow(n * M_PI, 2));
            gcirc[n] = gcirc[L - n] = (double)(float)(scale * (double)ramp);
        }
        for (int k = 0; k < L; k++) {
            double acc = gcirc[0];
            for (int n = 1; n <= g->nu - 1; n += 2) acc += 2.0 * gcirc[n] * ctab[(int)(((long long)k * n) & (L - 1))];
            spec[k] = (float)(acc / L);
        }
    }
    if (vc.size() > g_fdk.vc_cap) { cudaFree(g_fdk.d_vc); MONTE_CUDA(cudaMalloc(&g_fdk.d_vc, vc.size() * sizeof(ViewConst))); g_fdk.vc_cap = vc.size(); }
    if (wtab.size() > g_fdk.wtab_cap) { cudaFree(g_fdk.d_wtab); MONTE_CUDA(cudaMalloc(&g_fdk.d_wtab, wtab.size() * sizeof(double))); g_fdk.wtab_cap = wtab.size(); }
    if (taps.size() > g_fdk.taps_cap) { cudaFree(g_fdk.d_taps); MONTE_CUDA(cudaMalloc(&g_fdk.d_taps, taps.size() * sizeof(float))); g_fdk.taps_cap = taps.size(); }
    MONTE_CUDA(cudaMemcpyAsync(g_fdk.d_vc, vc.data(), vc.size() * sizeof(ViewConst), cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemcpyAsync(g_fdk.d_wtab, wtab.data(), wtab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemcpyAsync(g_fdk.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    if (fft_len) {
        if ((size_t)fft_len > g_fdk.fft_cap) {
            cudaFree(g_fdk.d_tw); cudaFree(g_fdk.d_spec);
            MONTE_CUDA(cudaMalloc(&g_fdk.d_tw, fft_len * sizeof(cpx)));
            MONTE_CUDA(cudaMalloc(&g_fdk.d_spec, fft_len * sizeof(float)));
            g_fdk.fft_cap = fft_len;
        }
        MONTE_CUDA(cudaMemcpyAsync(g_fdk.d_tw, tw.data(), fft_len * sizeof(cpx), cudaMemcpyHostToDevice, st));
        MONTE_CUDA(cudaMemcpyAsync(g_fdk.d_spec, spec.data(), fft_len * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    MONTE_CUDA(cudaStreamSynchronize(st));   // host vectors go out of scope
    g_fdk.g = key;
    g_fdk.noff = noff; g_fdk.tap_len = tap_len; g_fdk.nu_pad = nu_pad; g_fdk.in_pitch = in_pitch;
    g_fdk.fft_len = fft_len;
    g_fdk.valid = true;
    return MONTE_OK;
}

static size_t filtered_pitch(const monte_fdk_geom *g) { return (size_t)((g->nu + 1 + 3) / 4) * 4; }

}  // namespace monte

using namespace monte;

extern "C" {

size_t monte_gpu_fdk_filtered_pitch(const monte_fdk_geom *g) { return g ? filtered_pitch(g) : 0; }
size_t monte_gpu_fdk_filtered_elems(const monte_fdk_geom *g) {
    return g ? ((size_t)g->n_views * g->nv + 2) * filtered_pitch(g) : 0;
}

int monte_gpu_fdk_filter_dev(const monte_fdk_geom *g, const float *d_map, int view_begin, int view_end,
                             float *d_filtered_padded, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(d_map && d_filtered_padded, "fdk_filter: NULL device pointer");
    MONTE_ARG(0 <= view_begin && view_begin <= view_end && view_end <= g->n_views, "fdk_filter: bad view range");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = fdk_prepare(g, st)) return rc;
    if (view_begin == view_end) return MONTE_OK;
    // wide detectors: FFT filter (MONTE_FDK_FILTER=direct|fft overrides; default fft from nu > 256)
    bool use_fft = g_fdk.fft_len != 0 && g->nu > 256;
    if (const char *e = getenv("MONTE_FDK_FILTER")) {
        if (!strcmp(e, "direct")) use_fft = false;
        else if (!strcmp(e, "fft")) { MONTE_ARG(g_fdk.fft_len != 0, "fdk_filter: nu=%d too wide for the FFT filter", g->nu); use_fft = true; }
    }
    if (use_fft) {
        FftFilterParams q;
        q.map = d_map; q.wtab = g_fdk.d_wtab; q.tw = g_fdk.d_tw; q.spec = g_fdk.d_spec; q.out = d_filtered_padded;
        q.nu = g->nu; q.nv = g->nv; q.pitch = (int)filtered_pitch(g); q.view_begin = view_begin;
        const int L = g_fdk.fft_len;
        const size_t smem = fft_filter_smem(L);
        dim3 grid(ceil_div(g->nv, 2 * FFT_PAIRS), view_end - view_begin);
        if (!g_fft_attr_set) {
            MONTE_CUDA(cudaFuncSetAttribute(fdk_weight_filter_fft_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_filter_smem(1024)));
            MONTE_CUDA(cudaFuncSetAttribute(fdk_weight_filter_fft_kernel<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_filter_smem(2048)));
            MONTE_CUDA(cudaFuncSetAttribute(fdk_weight_filter_fft_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_filter_smem(4096)));
            g_fft_attr_set = true;
        }
        if (L == 1024) fdk_weight_filter_fft_kernel<1024> MONTE_CFG(grid, 128, smem, st)(q);
        else if (L == 2048) fdk_weight_filter_fft_kernel<2048> MONTE_CFG(grid, 256, smem, st)(q);
        else fdk_weight_filter_fft_kernel<4096> MONTE_CFG(grid, 512, smem, st)(q);
        MONTE_CUDA(cudaGetLastError());
        return MONTE_OK;
    }
    FilterParams p;
    p.map = d_map; p.wtab = g_fdk.d_wtab; p.taps = g_fdk.d_taps; p.out = d_filtered_padded;
    p.nu = g->nu; p.nv = g->nv; p.pitch = (int)filtered_pitch(g);
    p.view_begin = view_begin; p.nu_pad = g_fdk.nu_pad; p.in_pitch = g_fdk.in_pitch;
    p.noff = g_fdk.noff; p.tap_len = g_fdk.tap_len;
    const bool textbook = g->weight_mode == MONTE_FDK_TEXTBOOK;
    p.center = (float)((textbook ? g->dsd / (g->dso * g->du) : g->filter_scale) * 0.25);
    const size_t smem = ((size_t)2 * FT_D * p.in_pitch + p.tap_len) * sizeof(float);
    MONTE_ARG(smem <= 227 * 1024, "fdk_filter: nu=%d needs %zu B of shared memory (> 227 KB)", g->nu, smem);
    if (smem > g_filter_smem_set) {
        MONTE_CUDA(cudaFuncSetAttribute(fdk_weight_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_filter_smem_set = smem;
    }
    dim3 grid(ceil_div(g->nv, FT_D), view_end - view_begin);
    fdk_weight_filter_kernel MONTE_CFG(grid, FT_THREADS, smem, st)(p);
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

int monte_gpu_fdk_pad_dev(const monte_fdk_geom *g, float *d_filtered_padded, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    const int rows = g->n_views * g->nv;
    fdk_pad_kernel MONTE_CFG(ceil_div(rows + 2, 256), 256, 0, (cudaStream_t)stream)(d_filtered_padded, rows, g->nu, (int)filtered_pitch(g), 0, rows + 2);
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

int monte_gpu_fdk_pad_views_dev(const monte_fdk_geom *g, float *d_filtered_padded, int view_lo, int view_hi, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(0 <= view_lo && view_lo <= view_hi && view_hi <= g->n_views, "fdk_pad_views: bad view range");
    if (view_lo == view_hi) return MONTE_OK;
    const int rows = g->n_views * g->nv;
    const int r0 = view_lo * g->nv, r1 = view_hi == g->n_views ? rows + 2 : min(view_hi * g->nv + 2, rows + 2);
    fdk_pad_kernel MONTE_CFG(ceil_div(r1 - r0, 256), 256, 0, (cudaStream_t)stream)(d_filtered_padded, rows, g->nu, (int)filtered_pitch(g), r0, r1);
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

int monte_gpu_fdk_unpad_dev(const monte_fdk_geom *g, const float *d_filtered_padded, float *d_dense, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    const size_t rows = (size_t)g->n_views * g->nv, n = rows * g->nu;
    fdk_unpad_kernel MONTE_CFG((unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream)(d_filtered_padded, d_dense, rows, g->nu, (int)filtered_pitch(g));
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

// Backproject views [view_lo, view_hi) into z-slices [z_lo, z_hi).  continue_sum: the slab already
// holds the fp32 partial sums of the earlier views (they are reloaded exactly), else it is zeroed.
// axial detector rows [b_lo, b_hi) of every view that the slab [z_lo, z_hi) can reach: x = (half_v - k Z) / dv
// with k between the magnifications of the farthest and the nearest voxel; +-2 rows of slack, +1 for the
// pair's partner.  b_hi <= b_lo: the slab projects above or below the detector in every view.
static void slab_band(const monte_fdk_geom *g, int z_lo, int z_hi, int &b_lo, int &b_hi) {
    const double r = 0.5 * g->vox * sqrt((double)g->nx * g->nx + (double)g->ny * g->ny) +
                     fmax(fabs(g->x0 + 0.5 * g->vox * g->nx), fabs(g->y0 - 0.5 * g->vox * g->ny));
    const double kmin = g->dsd / (g->dso + r), kmax = g->dsd / fmax(g->dso - r, 0.05 * g->dso);
    const double Za = g->z0 - g->vox * z_lo, Zb = g->z0 - g->vox * (z_hi - 1);
    double xmin = 1e30, xmax = -1e30;
    for (int i = 0; i < 4; i++) {
        const double x = (g->half_v - ((i & 1) ? kmax : kmin) * ((i & 2) ? Zb : Za)) / g->dv;
        xmin = fmin(xmin, x); xmax = fmax(xmax, x);
    }
    b_lo = (int)floor(xmin) - 2; b_hi = (int)ceil(xmax) + 4;
    if (b_lo < 0) b_lo = 0;
    if (b_hi > g->nv) b_hi = g->nv;       // rows nv, nv+1 of a view are rows 0, 1 of the next one
}

// src != nullptr: the filtered rows live on the devices that filtered them (d_filtered_padded is not read)
static int backproject_views(const monte_fdk_geom *g, const float *d_filtered_padded, int z_lo, int z_hi,
                             float *d_vol_slab, cudaStream_t st, int view_lo, int view_hi, bool continue_sum,
                             const PairSrc *src = nullptr, int *zs_pending = nullptr) {
    // zs_pending (nullable): the caller feeds a thin slab chunk by chunk (fdk_multi) and joins the z-block streams itself
    // at the end (bp_join_zstreams): this call then leaves its launches on them -- *zs_pending = streams in use -- so that
    // the tail of one chunk's launches is filled by the next chunk's
    if (int rc = fdk_prepare(g, st)) return rc;
    if (z_lo == z_hi) return MONTE_OK;
    // everything outside the ROI is zero (the reference callocs the volume, bp3d20.cpp:32)
    if (!continue_sum) MONTE_CUDA(cudaMemsetAsync(d_vol_slab, 0, (size_t)(z_hi - z_lo) * g->ny * g->nx * sizeof(float), st));
    if (g->s_begin == g->s_end || g->t_begin == g->t_end || view_lo >= view_hi) return MONTE_OK;
    const bool textbook = g->weight_mode == MONTE_FDK_TEXTBOOK;
    BpParams p;
    p.vc = g_fdk.d_vc; p.vol = d_vol_slab;
    p.n_views = g->n_views; p.nu = g->nu; p.nv = g->nv; p.pitch = (int)filtered_pitch(g);
    p.nx = g->nx; p.ny = g->ny;
    p.s_begin = g->s_begin; p.s_end = g->s_end; p.t_begin = g->t_begin; p.t_end = g->t_end;
    p.z_lo = z_lo; p.z_hi = z_hi; p.roi_z_begin = g->z_begin; p.roi_z_end = g->z_end;
    p.x0 = (float)g->x0; p.y0 = (float)g->y0; p.z0 = (float)g->z0; p.vox = (float)g->vox;
    p.dso = (float)g->dso; p.dsd = (float)g->dsd;
    p.half_u = (float)g->half_u; p.half_v = (float)g->half_v;
    p.inv_du = (float)(1.0 / g->du); p.inv_dv = (float)(1.0 / g->dv);
    p.wd = (float)g->weight_dist; p.wd2 = (float)(g->weight_dist * g->weight_dist);
    const double dbeta = (double)(float)g->angle_step_deg * 2 * M_PI / 360;
    p.out_fac = (float)(textbook ? 0.5 * dbeta : dbeta * g->out_scale * g->out_scale2);
    p.eps_u = (float)(g->half_u * 2e-5); p.eps_v = (float)(g->half_v * 2e-5);
    p.eps_ts = (float)(2e-5 * (fabs(g->x0) + fabs(g->y0) + g->vox * (g->nx + g->ny)));
    {   // farthest a voxel column can be from the rotation axis -> smallest magnification
        const double ax = fmax(fabs(g->x0), fabs(g->x0 + g->vox * (g->nx - 1))), ay = fmax(fabs(g->y0), fabs(g->y0 - g->vox * (g->ny - 1)));
        p.kmin = (float)(0.99 * g->dsd / (g->dso + sqrt(ax * ax + ay * ay)));
    }
    p.textbook = textbook; p.coord_mode = g->coord_mode;
    p.mask_cs = g->mask_cs; p.mask_ct = g->mask_ct; p.mask_cz = g->mask_cz; p.mask_r2 = g->mask_r2;
    p.x0d = g->x0; p.y0d = g->y0; p.z0d = g->z0; p.voxd = g->vox; p.dsdd = g->dsd;
    p.half_ud = g->half_u; p.half_vd = g->half_v; p.inv_dud = 1.0 / g->du; p.inv_dvd = 1.0 / g->dv;
    p.nu_half = g->nu / 2.; p.nv_half = g->nv / 2.;
    dim3 block(BP_TX, BP_TY);
    // tuning knob (default = the fastest measured variant): MONTE_BP_VARIANT=0..3
    int variant;                                                     // (read per call: tests and A/B scripts flip it)
    { const char *e = getenv("MONTE_BP_VARIANT"); variant = e ? atoi(e) : BP_DEFAULT_VARIANT; }
#define BP_LAUNCH(ZT, ZB, MINB)                                                                          \
    do {                                                                                                 \
        dim3 grid(ceil_div(g->s_end - g->s_begin, BP_TX), ceil_div(g->t_end - g->t_begin, BP_TY),        \
                  ceil_div(z_hi - (z_lo / ZT) * ZT, ZT));                                                \
        fdk_backproject_kernel<ZT, ZB, MINB> MONTE_CFG(grid, block, 0, st)(p);                                 \
    } while (0)
    // Views are streamed in chunks so that the detector rows a wave of CTAs touches stay L2-resident
    // (all 720 views of one z-block are ~200 MB; CTAs drifting apart in view index thrash the L2).
    // fp32 partial sums are stored and reloaded exactly, so chunking does not change the result.
    int vchunk;
    {
        static int env_chunk = -2;
        if (env_chunk == -2) { const char *e = getenv("MONTE_BP_VCHUNK"); env_chunk = e ? atoi(e) : -1; }
        if (env_chunk >= 0) vchunk = env_chunk > 0 ? env_chunk : 1 << 30;
        else {   // rows of one view that a 32-slice z-block projects onto, times the row size, against ~64 MB of L2
            const double r = 0.5 * g->vox * sqrt((double)g->nx * g->nx + (double)g->ny * g->ny);
            const double mag = g->dsd / fmax(g->dso - r, 0.1 * g->dso);
            const double band_rows = 32.0 * g->vox / g->dv * mag + 8.0;
            const double per_view = band_rows * (double)p.pitch * sizeof(float);
            vchunk = (int)fmax(30.0, 64.0 * 1024 * 1024 / per_view);
        }
    }
    // vertical row pairs of the views of this call (+ the rows the last view reaches into), converted view chunk by view
    // chunk on a stream of their own (below): chunk c + 1 is converted -- between devices: fetched over NVLink -- while
    // chunk c is backprojected
    const int rows_total = g->n_views * g->nv + 2;
    float2 *d_pairs = (float2 *)scratch(8, (size_t)rows_total * p.pitch * sizeof(float2) + 65536);   // (+ slack: staged tile rows may end past the last row)
    if (!d_pairs) return MONTE_E_NOMEM;
    int b_lo, b_hi;
    slab_band(g, z_lo, z_hi, b_lo, b_hi);
    // the whole slab projects above or below the detector in every view (bp3d20.cpp:116 skips every
    // voxel of it): its voxels keep the zeros / partial sums they have
    if (b_hi <= b_lo) return MONTE_OK;
    // rows 0..3 of every view are always paired too: they are what the previous view reaches past its end -- also
    // those of the view after the last one of a chunk (and only those of it: the rest of that view may not be
    // filtered yet when a multi-device caller feeds the views chunk by chunk)
    // (one kernel for both sources: the local padded rows as a single segment -- the pad rules it applies on the fly are
    // the ones fdk_pad_kernel wrote into them -- or the peers' rows over NVLink)
    PairSrc local_src;
    if (!src) { local_src.n = 1; local_src.base[0] = d_filtered_padded; local_src.v_end[0] = g->n_views; }
    const PairSrc &psrc = src ? *src : local_src;
    auto pair_views = [&](int v_a, int v_b, cudaStream_t ss) -> int {
        const int a_hi = b_lo > 0 ? (b_lo < 4 ? b_lo : 4) : 0;
        const int x_hi = v_b < g->n_views ? 4 : 2;                  // rows of the next view / the two zero rows after the last view
        const dim3 grid(1 + ceil_div(b_hi - b_lo, PAIR_RPT), ceil_div(p.pitch, 128), v_b - v_a + 1);
        int seg0 = 0;
        while (seg0 < psrc.n - 1 && v_a >= psrc.v_end[seg0]) seg0++;
        fdk_pair_gather_kernel MONTE_CFG(grid, 128, 0, ss)(psrc, d_pairs, v_a, v_b - v_a, g->nv, g->nu, a_hi, b_lo, b_hi, x_hi, rows_total - 2, p.pitch, seg0);
        MONTE_CUDA(cudaGetLastError());
        return MONTE_OK;
    };
    // A thin slab (one of N multi-GPU slabs: 3..7 z-blocks of 16 slices) makes every view-chunk launch a handful of
    // waves (C3 on 8 GPUs: 3072 CTAs on 592 resident slots = 5.2), and the launches of one stream do not overlap: each
    // ends in a partly filled wave, ~10 % of the slab's time.  The z-blocks are independent, so each gets its own
    // stream: its view chunks stay in order, and the tail of one z-block's launch is filled with CTAs of the next
    // one's.  Launches are issued chunk-major, so the z-blocks walk through the views together (L2-sized chunks).
    const int zb_first = z_lo / 16, zb_end = ceil_div(z_hi, 16);
    bool zsplit = variant == 0 && zb_end - zb_first >= 2 && zb_end - zb_first <= 8 && (view_hi - view_lo > vchunk || zs_pending != nullptr);
#ifdef MONTE_EMU
    zsplit = false;
#endif
    { const char *e = getenv("MONTE_BP_ZSTREAMS"); if (e && atoi(e) == 0) zsplit = false; }   // (read per call: the test flips it)
    FdkDevState &zds = g_fdk_state.get();
    if (zsplit) {
        if (!zds.ev_zfork) {
            MONTE_CUDA(cudaEventCreateWithFlags(&zds.ev_zfork, cudaEventDisableTiming));
            for (int i = 0; i < 8; i++) {
                MONTE_CUDA(cudaStreamCreateWithFlags(&zds.zs[i], cudaStreamNonBlocking));
                MONTE_CUDA(cudaEventCreateWithFlags(&zds.ev_zjoin[i], cudaEventDisableTiming));
            }
        }
        MONTE_CUDA(cudaEventRecord(zds.ev_zfork, st));
        for (int i = 0; i < zb_end - zb_first; i++) MONTE_CUDA(cudaStreamWaitEvent(zds.zs[i], zds.ev_zfork, 0));
    }
    int n_vchunks = ceil_div(view_hi - view_lo, vchunk);
    bool psplit = n_vchunks >= 2 && n_vchunks < 32;
#ifdef MONTE_EMU
    psplit = false;
#endif
    // With the conversion on its own stream only the FIRST chunk's conversion is exposed: it is made a quarter chunk
    // (between devices it is an NVLink gather: 0.65 ms of a 9 ms slab at C3 on 8 GPUs).
    // (local rows: the conversion of a chunk is 0.25 ms and an extra launch costs more than that)
    const int first_chunk = psplit && src ? (vchunk / 4 > 8 ? vchunk / 4 : 8) : vchunk;
    if (first_chunk != vchunk) n_vchunks = 1 + ceil_div(view_hi - view_lo - first_chunk > 0 ? view_hi - view_lo - first_chunk : 0, vchunk);
    if (psplit) {
        if (!zds.ps) {
            MONTE_CUDA(cudaStreamCreateWithFlags(&zds.ps, cudaStreamNonBlocking));
            MONTE_CUDA(cudaEventCreateWithFlags(&zds.ev_pfork, cudaEventDisableTiming));
            for (int i = 0; i < 32; i++) MONTE_CUDA(cudaEventCreateWithFlags(&zds.ev_pair[i], cudaEventDisableTiming));
        }
        MONTE_CUDA(cudaEventRecord(zds.ev_pfork, st));
        MONTE_CUDA(cudaStreamWaitEvent(zds.ps, zds.ev_pfork, 0));
    }
    int ci = 0;
    for (int vb = view_lo, vstep = first_chunk; vb < view_hi; vb += vstep, vstep = vchunk, ci++) {
    // a chunk is presented to the kernel as a shorter scan: shifted view constants and rows
    p.n_views = vb + vstep < view_hi ? vstep : view_hi - vb;
    if (int rc = pair_views(vb, vb + p.n_views, psplit ? zds.ps : st)) return rc;
    if (psplit) {                                                  // the chunk's backprojection waits for its pairs only
        MONTE_CUDA(cudaEventRecord(zds.ev_pair[ci], zds.ps));
        if (zsplit) { for (int i = 0; i < zb_end - zb_first; i++) MONTE_CUDA(cudaStreamWaitEvent(zds.zs[i], zds.ev_pair[ci], 0)); }
        else MONTE_CUDA(cudaStreamWaitEvent(st, zds.ev_pair[ci], 0));
    } else if (zsplit) {                                           // pairs converted on st: the z-block streams wait for them
        MONTE_CUDA(cudaEventRecord(zds.ev_zfork, st));
        for (int i = 0; i < zb_end - zb_first; i++) MONTE_CUDA(cudaStreamWaitEvent(zds.zs[i], zds.ev_zfork, 0));
    }
    p.vc = g_fdk.d_vc + vb;
    p.pairs = d_pairs + (size_t)vb * g->nv * p.pitch; p.accumulate = continue_sum || vb > view_lo;
    if (zsplit) {
        for (int zb = zb_first; zb < zb_end; zb++) {
            BpParams q = p;
            q.z_lo = zb * 16 > z_lo ? zb * 16 : z_lo; q.z_hi = (zb + 1) * 16 < z_hi ? (zb + 1) * 16 : z_hi;
            q.vol = d_vol_slab + (size_t)(q.z_lo - z_lo) * g->ny * g->nx;
            dim3 grid(ceil_div(g->s_end - g->s_begin, BP_TX), ceil_div(g->t_end - g->t_begin, BP_TY), 1);
            fdk_backproject_kernel<16, 8, 4> MONTE_CFG(grid, block, 0, zds.zs[zb - zb_first])(q);
        }
        continue;
    }
    switch (variant) {
#ifndef MONTE_EMU
        case 10: case 11: {
            // footprint staged in shared memory by TMA (fdk_backproject_smem_kernel): box = the largest rectangle of row
            // pairs a 16 x 16 x 16 brick can project onto, from the geometry
            const double rf = 0.5 * g->vox * sqrt((double)g->nx * g->nx + (double)g->ny * g->ny) +
                              fmax(fabs(g->x0 + 0.5 * g->vox * g->nx), fabs(g->y0 - 0.5 * g->vox * g->ny));
            const double kmax = g->dsd / fmax(g->dso - rf, 0.1 * g->dso);
            const double diag = (BPS_T - 1) * sqrt(2.0);
            int C = (int)ceil(diag * g->vox * kmax / g->du + 5.0);
            C = (C + 1) & ~1;
            const double zmax = fmax(fabs(g->z0), fabs(g->z0 - g->vox * (g->nz - 1)));
            const double dk = kmax * (diag * g->vox) / fmax(g->dso - rf, 0.1 * g->dso);
            const int R = (int)ceil(15.0 * g->vox * kmax / g->dv + zmax / g->dv * dk + 4.0);
            const int stage_elems = (R * C + 15) & ~15;                                   // stages start 128-byte aligned
            const size_t smem = (size_t)BPS_STAGES * stage_elems * sizeof(float2) + 2 * BPS_STAGES * sizeof(unsigned long long) + BPS_STAGES * sizeof(int2);
            if (C >= 8 && C <= 128 && R <= 256 && smem <= 100 * 1024) {
                FdkDevState &ds = g_fdk_state.get();
                auto fn = variant == 10 ? fdk_backproject_smem_kernel<16, 8, 4> : fdk_backproject_smem_kernel<16, 8, 3>;
                if (smem > ds.bps_smem_set[variant - 10]) {
                    MONTE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    ds.bps_smem_set[variant - 10] = smem;
                }
                CUtensorMap tmap;
                if (int rc = make_pairs_tensor_map(&tmap, p.pairs, (size_t)rows_total - (size_t)vb * g->nv, p.pitch, C, R)) return rc;
                dim3 grid(ceil_div(g->s_end - g->s_begin, BPS_T), ceil_div(g->t_end - g->t_begin, BPS_T), ceil_div(z_hi - (z_lo / 16) * 16, 16));
                fn<<<grid, BPS_T * BPS_T, smem, st>>>(p, tmap, C, R, stage_elems);
                break;
            }
        }   // (tile too large for shared memory: the L1 kernel)
        // fallthrough
#endif
        case 0: default: BP_LAUNCH(16, 8, 4); break;    // 64 registers (4 bytes spilled), 4 CTAs/SM = 32 warps: 56.2 ms -- the gathers'
                                                // latency (long scoreboard, the top stall) is hidden by more resident warps
        case 1: BP_LAUNCH(32, 8, 2); break;     // 32 slices per thread, 2 CTAs/SM: 74.8 ms at C3 (before the 12-instruction update)
        case 2: BP_LAUNCH(16, 8, 3); break;     // 80 registers, 3 CTAs/SM: 61.9 ms
        case 3: BP_LAUNCH(16, 16, 3); break;    // gather batches of 16: 58.6 ms
    }
    }
#undef BP_LAUNCH
    MONTE_CUDA(cudaGetLastError());
    if (zsplit && zs_pending) *zs_pending = zb_end - zb_first;     // (joined by the caller)
    else if (zsplit)
        for (int i = 0; i < zb_end - zb_first; i++) {
            MONTE_CUDA(cudaEventRecord(zds.ev_zjoin[i], zds.zs[i]));
            MONTE_CUDA(cudaStreamWaitEvent(st, zds.ev_zjoin[i], 0));
        }
    return MONTE_OK;
}

// st waits for the first n z-block streams of the current device (the deferred join of backproject_views)
static int bp_join_zstreams(cudaStream_t st, int n) {
    FdkDevState &zds = g_fdk_state.get();
    for (int i = 0; i < n; i++) {
        MONTE_CUDA(cudaEventRecord(zds.ev_zjoin[i], zds.zs[i]));
        MONTE_CUDA(cudaStreamWaitEvent(st, zds.ev_zjoin[i], 0));
    }
    return MONTE_OK;
}

int monte_gpu_fdk_slab_rows(const monte_fdk_geom *g, int z_lo, int z_hi, int *row_lo, int *row_hi) {
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(row_lo && row_hi && 0 <= z_lo && z_lo <= z_hi && z_hi <= g->nz, "fdk_slab_rows: bad argument");
    int a = 0, b = 0;
    if (z_hi > z_lo) slab_band(g, z_lo, z_hi, a, b);
    if (b < a) b = a;
    *row_lo = a; *row_hi = b;
    return MONTE_OK;
}

int monte_gpu_fdk_backproject_dev(const monte_fdk_geom *g, const float *d_filtered_padded, int z_lo, int z_hi,
                                  float *d_vol_slab, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(d_filtered_padded && d_vol_slab, "fdk_backproject: NULL device pointer");
    MONTE_ARG(0 <= z_lo && z_lo <= z_hi && z_hi <= g->nz, "fdk_backproject: bad z range [%d,%d)", z_lo, z_hi);
    return backproject_views(g, d_filtered_padded, z_lo, z_hi, d_vol_slab, (cudaStream_t)stream, 0, g->n_views, false);
}

int monte_gpu_fdk_backproject_views_dev(const monte_fdk_geom *g, const float *d_filtered_padded, int z_lo, int z_hi,
                                        float *d_vol_slab, int view_lo, int view_hi, int continue_sum, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(d_filtered_padded && d_vol_slab, "fdk_backproject_views: NULL device pointer");
    MONTE_ARG(0 <= z_lo && z_lo <= z_hi && z_hi <= g->nz, "fdk_backproject_views: bad z range [%d,%d)", z_lo, z_hi);
    MONTE_ARG(0 <= view_lo && view_lo <= view_hi && view_hi <= g->n_views, "fdk_backproject_views: bad view range");
    return backproject_views(g, d_filtered_padded, z_lo, z_hi, d_vol_slab, (cudaStream_t)stream, view_lo, view_hi, continue_sum != 0);
}

int monte_gpu_fdk_backproject_peers_dev(const monte_fdk_geom *g, int n_seg, const float *const *seg_base, const int *seg_v_end,
                                        int z_lo, int z_hi, float *d_vol_slab, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(seg_base && seg_v_end && d_vol_slab && n_seg >= 1 && n_seg <= PAIR_MAX_SEG, "fdk_backproject_peers: bad argument (n_seg = %d, at most %d)", n_seg, PAIR_MAX_SEG);
    MONTE_ARG(0 <= z_lo && z_lo <= z_hi && z_hi <= g->nz, "fdk_backproject_peers: bad z range");
    PairSrc src;
    src.n = n_seg;
    int prev = 0;
    for (int o = 0; o < n_seg; o++) {
        MONTE_ARG(seg_base[o] && seg_v_end[o] >= prev, "fdk_backproject_peers: segments must ascend in views and have a base");
        src.base[o] = seg_base[o]; src.v_end[o] = seg_v_end[o];
        prev = seg_v_end[o];
    }
    MONTE_ARG(prev == g->n_views, "fdk_backproject_peers: the segments cover %d of %d views", prev, g->n_views);
    return backproject_views(g, nullptr, z_lo, z_hi, d_vol_slab, (cudaStream_t)stream, 0, g->n_views, false, &src);
}

// ---- CUDA IPC (one process per GPU): the containing allocation of a pointer is found through the driver entry point
// cuMemGetAddressRange (libmonte_gpu does not link libcuda); opened mappings are remembered so that close() can unmap
struct IpcMap { void *base; void *user; };
static std::vector<IpcMap> g_ipc_maps;
int monte_gpu_ipc_export(const void *d_ptr, unsigned char handle[MONTE_IPC_HANDLE_BYTES], uint64_t *offset) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(d_ptr && handle && offset, "ipc_export: NULL argument");
#ifdef MONTE_EMU
    memset(handle, 0, MONTE_IPC_HANDLE_BYTES);
    memcpy(handle, &d_ptr, sizeof(d_ptr));                          // (one address space: the "handle" is the pointer)
    *offset = 0;
    return MONTE_OK;
#else
    static_assert(sizeof(cudaIpcMemHandle_t) == MONTE_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static range_fn get_range = nullptr;
    if (!get_range) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        MONTE_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &qr));
        if (!sym || qr != cudaDriverEntryPointSuccess) { set_error("cuMemGetAddressRange is not available in this driver"); return MONTE_E_CUDA; }
        get_range = (range_fn)sym;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (CUdeviceptr)d_ptr) != CUDA_SUCCESS) { set_error("ipc_export: %p is not inside a device allocation", d_ptr); return MONTE_E_ARG; }
    cudaIpcMemHandle_t h;
    MONTE_CUDA(cudaIpcGetMemHandle(&h, (void *)base));
    memcpy(handle, &h, sizeof(h));
    *offset = (uint64_t)((CUdeviceptr)d_ptr - base);
    return MONTE_OK;
#endif
}
int monte_gpu_ipc_open(const unsigned char handle[MONTE_IPC_HANDLE_BYTES], uint64_t offset, void **d_ptr) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(handle && d_ptr, "ipc_open: NULL argument");
    void *base = nullptr;
#ifdef MONTE_EMU
    memcpy(&base, handle, sizeof(base));
#else
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    MONTE_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
#endif
    *d_ptr = (char *)base + offset;
    g_ipc_maps.push_back({base, *d_ptr});
    return MONTE_OK;
}
int monte_gpu_ipc_close(void *d_ptr) {
    for (size_t i = 0; i < g_ipc_maps.size(); i++)
        if (g_ipc_maps[i].user == d_ptr) {
#ifndef MONTE_EMU
            const cudaError_t e = cudaIpcCloseMemHandle(g_ipc_maps[i].base);
            if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle", __FILE__, __LINE__);
#endif
            g_ipc_maps.erase(g_ipc_maps.begin() + i);
            return MONTE_OK;
        }
    set_error("ipc_close: %p was not returned by monte_gpu_ipc_open", d_ptr);
    return MONTE_E_ARG;
}

int monte_gpu_fdk_transpose_dev(const monte_fdk_geom *g, const float *d_vol_xy, float *d_vol_zy, void *stream) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    dim3 grid(ceil_div(g->nx, 32), ceil_div(g->nz, 32), g->ny);
    fdk_transpose_kernel MONTE_CFG(grid, dim3(32, 8), 0, (cudaStream_t)stream)(d_vol_xy, d_vol_zy, g->nx, g->ny, g->nz);
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

// ------------------------------------------------------------------------------------------
// multi-device reconstruction: filter sharded by views, backprojection by z-slabs of equal work
// ------------------------------------------------------------------------------------------
// Relative backprojection cost of every z-slice: at wide cone angles the end slices see the detector in few views
// or none.  cost = overhead + f (1 + penalty [f below its maximum]), f = fraction of the slice's (voxel, view) pairs
// that project onto the detector (bp3d20.cpp:116 skips the others), sampled on a coarse (s, t, view) grid with the
// reference's projection formulas (:99-113); `overhead` = per-column work done while any view sees the slice,
// `penalty` = the slower general path of columns near the detector edge.  Calibrated on one B200 at C3
// (scripts/fdk_slab_cost.py); the same model as monte_b200/dist.py:fdk_slice_cost.
static void fdk_slice_cost(const monte_fdk_geom *g, std::vector<double> &cost) {
    const double overhead = 0.32, penalty = 0.45;
    const int stride = 8;
    std::vector<float> ks;
    size_t total = 0;
    for (int t = 0; t < g->ny; t += stride)
        for (int s = 0; s < g->nx; s += stride)
            for (int v = 0; v < g->n_views; v += stride) {
                const double X = g->x0 + g->vox * s, Y = g->y0 - g->vox * t;
                const double beta = M_PI * (g->angle0_deg + g->angle_step_deg * v) / 180;
                const double k = g->dsd / (X * cos(beta) + Y * sin(beta) + g->dso);
                total++;
                if (fabs(k * (-X * sin(beta) + Y * cos(beta))) <= g->half_u) ks.push_back((float)k);
            }
    std::sort(ks.begin(), ks.end());
    cost.assign(g->nz, 0.0);
    std::vector<double> frac(g->nz);
    double fmaxv = 0;
    for (int z = 0; z < g->nz; z++) {
        const double Z = fmax(fabs(g->z0 - g->vox * z), 1e-12);
        // |k Z| <= half_v  <=>  k <= half_v / |Z|
        frac[z] = (double)(std::upper_bound(ks.begin(), ks.end(), (float)(g->half_v / Z)) - ks.begin()) / (double)(total ? total : 1);
        fmaxv = fmax(fmaxv, frac[z]);
    }
    for (int z = 0; z < g->nz; z++) {
        const bool partial = frac[z] > 0 && frac[z] < 0.97 * fmaxv;
        cost[z] = (frac[z] > 0 ? overhead : 0.02) + frac[z] * (1.0 + (partial ? penalty : 0.0));
    }
}

// contiguous cuts of range(n) into `world` pieces of nearly equal cost, moved to multiples of `align` (the
// backprojector's z-block) where that keeps every piece non-empty; cuts[0] = 0 .. cuts[world] = n
static void balanced_cuts(const std::vector<double> &cost, int world, int align, int *cuts) {
    const int n = (int)cost.size();
    std::vector<double> acc(n + 1, 0.0);
    for (int i = 0; i < n; i++) acc[i + 1] = acc[i] + cost[i];
    cuts[0] = 0;
    for (int r = 1; r < world; r++) {
        const double target = acc[n] * r / world;
        const int lo = n >= world ? cuts[r - 1] + 1 : cuts[r - 1], hi = n >= world ? n - (world - r) : n;
        int k = lo;
        while (k < hi && acc[k] < target) k++;
        if (k > lo && target - acc[k - 1] < acc[k] - target) k--;
        k = k < lo ? lo : (k > hi ? hi : k);
        if (align > 1) {
            const int ka = (k + align / 2) / align * align;
            if (lo <= ka && ka <= hi) k = ka;
        }
        cuts[r] = k;
    }
    cuts[world] = n;
}

struct FdkMultiDev {              // what one device holds during a multi-device reconstruction
    int n_views = 0, z_lo = 0, z_hi = 0;                 // views it filters (all its chunks), its z-slab
    int off[8] = {0, 0, 0, 0, 0, 0, 0, 0};               // first local view of its k-th chunk
    float *d_map = nullptr, *d_filt = nullptr, *d_vol = nullptr;
    cudaEvent_t start = nullptr, filter_end = nullptr, bp_end = nullptr;   // timing events on this device
};

int monte_gpu_fdk_partition(const monte_fdk_geom *g, int n_parts, int *z_cuts) {
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(n_parts >= 1 && z_cuts, "fdk_partition: bad argument");
    std::vector<double> cost;
    fdk_slice_cost(g, cost);
    std::vector<int> cuts(n_parts + 1);
    balanced_cuts(cost, n_parts, 16, cuts.data());
    for (int i = 0; i <= n_parts; i++) z_cuts[i] = cuts[i];
    return MONTE_OK;
}

// monte_gpu_fdk on all bound devices.  The views are cut into C = N x n_up chunks in ascending order; chunk c
// belongs to device c mod N, which uploads it over its own PCIe link and filters it.  Every device backprojects ALL
// views into its own z-slab, chunk by chunk in ascending order, as soon as a chunk (and the one after it, whose first
// rows the last view of a chunk reaches into) has been filtered wherever it lives: the detector-row band the slab reads
// is loaded straight out of the owner's filtered rows (fdk_pair_gather_kernel: the exchange is fused into the pair
// conversion over NVLink peer loads, nothing is packed, copied or stored twice) and the fp32 partial sums of the slab
// are continued exactly.  Interleaving the chunks over the devices is what lets the reconstruction run while the
// later chunks are still on their way: all PCIe links are busy from the start AND the first views are complete after
// 1/n_up of the upload time (with contiguous view ranges nothing could start before device 0 had all of its views).
// Three streams per device: copy (H2D chunks, D2H slab), stream (filter), aux (pair gather + backprojection).
// Views are consumed in ascending order on every device, so the volume has the bits of the one-device call.
static int fdk_multi(const monte_fdk_geom *g, const float *map, float *filtered, float *vol_xy, float *vol_zy, monte_fdk_stats *stats) {
    const int nd = n_dev();
    if (!peers_ok()) {
        set_error("fdk: the %d bound devices lack mutual peer access (NVLink P2P), which the multi-device reconstruction needs", nd);
        return MONTE_E_NODEV;
    }
    const auto t_host0 = std::chrono::steady_clock::now();
    const size_t per_view = (size_t)g->nu * g->nv, slice = (size_t)g->nx * g->ny;
    const int pitch = (int)filtered_pitch(g);
    int cuts[MAX_DEV + 1];
    {
        static monte_fdk_geom key;                                  // the partition depends on the geometry only
        static int key_nd = 0, key_cuts[MAX_DEV + 1];
        const monte_fdk_geom cg = canon_geom(g);
        if (key_nd != nd || memcmp(&key, &cg, sizeof(cg)) != 0) {
            std::vector<double> cost;
            fdk_slice_cost(g, cost);
            balanced_cuts(cost, nd, 16, key_cuts);
            key = cg; key_nd = nd;
        }
        memcpy(cuts, key_cuts, sizeof(cuts));
    }
    // chunks of views: c = k * nd + i is the k-th chunk of device i
    // (as many chunks as the segment table holds, at most 8 per device, at least 4 views each: finer chunks start the
    // backprojection earlier and shorten what is left of it after the last upload; each costs ~4 launches per device)
    int n_up = PAIR_MAX_SEG / nd < 8 ? PAIR_MAX_SEG / nd : 8;
    while (n_up > 1 && g->n_views < 4 * nd * n_up) n_up--;
    { const char *e = getenv("MONTE_FDK_MULTI_CHUNKS"); if (e && atoi(e) >= 1 && atoi(e) <= 8 && atoi(e) * nd <= PAIR_MAX_SEG) n_up = atoi(e); }
    const int C = nd * n_up;
    int V[PAIR_MAX_SEG + 1];
    for (int c = 0; c <= C; c++) V[c] = (int)((long long)g->n_views * c / C);
    FdkMultiDev dv[MAX_DEV];
    PairSrc src;
    src.n = C;
    cudaEvent_t ev_f[PAIR_MAX_SEG] = {nullptr};                      // chunk c is filtered (recorded on its owner's stream)
    int rc = MONTE_OK, launches = 0;
    for (int i = 0; i < nd; i++) {
        int n = 0;
        for (int k = 0; k < n_up; k++) { dv[i].off[k] = n; n += V[k * nd + i + 1] - V[k * nd + i]; }
        dv[i].n_views = n; dv[i].z_lo = cuts[i]; dv[i].z_hi = cuts[i + 1];
    }
    // ---- on every device: upload | filter, chunk by chunk
    for (int i = 0; i < nd && rc == MONTE_OK; i++) {
        FdkMultiDev &d = dv[i];
        if ((rc = use_dev(i))) break;
        Context &c = ctx();
        cudaStream_t st = c.stream, cp = c.copy_stream;
        d.d_map = (float *)scratch(0, (size_t)(d.n_views ? d.n_views : 1) * per_view * sizeof(float));
        d.d_filt = (float *)scratch(1, ((size_t)d.n_views * g->nv + 2) * pitch * sizeof(float));
        d.d_vol = (float *)scratch(2, (size_t)(d.z_hi > d.z_lo ? d.z_hi - d.z_lo : 1) * slice * sizeof(float));
        if (!d.d_map || !d.d_filt || !d.d_vol) { rc = MONTE_E_NOMEM; break; }
        if ((rc = fdk_prepare(g, st))) break;                       // (synchronises on a miss: the tables are complete for every stream)
        FdkDevState &ds = g_fdk_state.get();
        if (!ds.ev_up[0]) {
            for (int k = 0; k < FDK_MAXC && rc == MONTE_OK; k++)
                if (cudaEventCreateWithFlags(&ds.ev_up[k], cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ds.ev_slab[k], cudaEventDisableTiming) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
            for (int k = 0; k < 6 && rc == MONTE_OK; k++)
                if (cudaEventCreate(&ds.ev_t[k]) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
            if (rc) break;
        }
        d.start = ds.ev_t[0]; d.filter_end = ds.ev_t[1]; d.bp_end = ds.ev_t[2];
        if (cudaEventRecord(d.start, st) != cudaSuccess || cudaStreamWaitEvent(cp, d.start, 0) != cudaSuccess ||
            cudaStreamWaitEvent(c.aux_stream, d.start, 0) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__); break; }
        for (int k = 0; k < n_up && rc == MONTE_OK; k++) {
            const int ch = k * nd + i, a = V[ch], b = V[ch + 1];
            // row R of the global padded-row layout, R in chunk ch, lives at base + R * pitch on its owner
            src.base[ch] = d.d_filt + ((long long)d.off[k] - a) * g->nv * pitch;
            src.v_end[ch] = b;
            if (b > a) {
                float *dm = d.d_map + (size_t)d.off[k] * per_view;
                if (cudaMemcpyAsync(dm, map + (size_t)a * per_view, (size_t)(b - a) * per_view * sizeof(float), cudaMemcpyHostToDevice, cp) != cudaSuccess ||
                    cudaEventRecord(ds.ev_up[k], cp) != cudaSuccess || cudaStreamWaitEvent(st, ds.ev_up[k], 0) != cudaSuccess) {
                    rc = cuda_fail(cudaGetLastError(), "upload of a view chunk", __FILE__, __LINE__); break;
                }
                // the filter addresses maps and rows by absolute view: hand it the virtual bases
                if ((rc = monte_gpu_fdk_filter_dev(g, dm - (size_t)a * per_view, a, b, const_cast<float *>(src.base[ch]), st))) break;
                launches++;
                if (filtered) {                                     // the filtered views, if wanted: dense rows through the (now free) map chunk
                    const size_t rows = (size_t)(b - a) * g->nv, n = rows * g->nu;
                    fdk_unpad_kernel MONTE_CFG((unsigned)((n + 255) / 256), 256, 0, st)(d.d_filt + (size_t)d.off[k] * g->nv * pitch, dm, rows, g->nu, pitch);
                    launches++;
                }
            }
            if (cudaEventCreateWithFlags(&ev_f[ch], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev_f[ch], st) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__);
        }
        if (rc) break;
        if (cudaEventRecord(d.filter_end, st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__);
    }
    // ---- on every device: chunk by chunk, gather the band from the chunk's owner + backproject into the own slab.
    // Every device's launches are issued by a host thread of its own: ~250 launches per device (pair conversion + one
    // backprojection launch per z-block, for each of up to 32 chunks) would otherwise queue behind each other in one
    // thread for ~10 ms -- as long as the uploads they are meant to hide under.
    bool first_bp[MAX_DEV];
    int zs_pending[MAX_DEV];                                        // z-block streams a device's chunk launches were left on
    int launches_dev[MAX_DEV], rc_dev[MAX_DEV];
    std::string err_dev[MAX_DEV];
    for (int i = 0; i < nd; i++) { first_bp[i] = true; zs_pending[i] = 0; launches_dev[i] = 0; rc_dev[i] = MONTE_OK; }
    auto device_chunks = [&](int i) {
        FdkMultiDev &d = dv[i];
        int r = MONTE_OK;
        do {
            if (d.z_hi <= d.z_lo) break;
            if ((r = use_dev(i))) break;
            cudaStream_t bp = ctx().aux_stream;
            int waited = -1;                                        // chunks [0, waited] are known to this stream
            for (int ch = 0; ch < C && r == MONTE_OK; ch++) {
                if (V[ch + 1] <= V[ch]) continue;
                int need = ch + 1;                                  // ... up to the next chunk that holds a view
                while (need < C - 1 && V[need + 1] <= V[need]) need++;
                if (need > C - 1) need = C - 1;
                if (g->nv < 8) need = C - 1;                        // (rows 0..3 of "the next view" then span several views)
                for (int w = waited + 1; w <= need && r == MONTE_OK; w++)   // (own chunks too: they are filtered on another stream)
                    if (cudaStreamWaitEvent(bp, ev_f[w], 0) != cudaSuccess) r = cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
                if (r) break;
                if (need > waited) waited = need;
                if ((r = backproject_views(g, nullptr, d.z_lo, d.z_hi, d.d_vol, bp, V[ch], V[ch + 1], !first_bp[i], &src, &zs_pending[i]))) break;
                first_bp[i] = false;
                launches_dev[i] += 2;
            }
        } while (0);
        rc_dev[i] = r;
        if (r) err_dev[i] = monte_gpu_last_error();                 // (the error text is thread-local)
    };
    if (rc == MONTE_OK) {
#ifdef MONTE_EMU
        for (int i = 0; i < nd; i++) device_chunks(i);              // (the emulation runs kernels on the calling thread)
#else
        std::thread th[MAX_DEV];
        for (int i = 1; i < nd; i++) th[i] = std::thread(device_chunks, i);
        device_chunks(0);
        for (int i = 1; i < nd; i++) th[i].join();
#endif
        for (int i = 0; i < nd; i++) {
            launches += launches_dev[i];
            if (rc_dev[i] && rc == MONTE_OK) { rc = rc_dev[i]; set_error("%s", err_dev[i].c_str()); }
        }
    }
    for (int i = 0; i < nd && rc == MONTE_OK; i++) {
        FdkMultiDev &d = dv[i];
        if ((rc = use_dev(i))) break;
        Context &c = ctx();
        cudaStream_t bp = c.aux_stream, cp = c.copy_stream;
        const int nz_i = d.z_hi - d.z_lo;
        if (zs_pending[i] && (rc = bp_join_zstreams(bp, zs_pending[i]))) break;
        if (nz_i > 0 && first_bp[i]) {                              // no view at all: the slab is zero
            if (cudaMemsetAsync(d.d_vol, 0, (size_t)nz_i * slice * sizeof(float), bp) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaMemsetAsync", __FILE__, __LINE__); break; }
        }
        if (cudaEventRecord(d.bp_end, bp) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__); break; }
        if (nz_i > 0) {                                             // the slab goes home
            if (cudaStreamWaitEvent(cp, d.bp_end, 0) != cudaSuccess ||
                cudaMemcpyAsync(vol_xy + (size_t)d.z_lo * slice, d.d_vol, (size_t)nz_i * slice * sizeof(float), cudaMemcpyDeviceToHost, cp) != cudaSuccess)
                { rc = cuda_fail(cudaGetLastError(), "download of a slab", __FILE__, __LINE__); break; }
        }
        if (vol_zy && nz_i > 0) {                                   // image_zy[s][t][z] (bp3d20.cpp:161): this slab's z-columns
            float *d_zy = (float *)scratch(3, (size_t)nz_i * slice * sizeof(float));
            if (!d_zy) { rc = MONTE_E_NOMEM; break; }
            dim3 grid(ceil_div(g->nx, 32), ceil_div(nz_i, 32), g->ny);
            fdk_transpose_kernel MONTE_CFG(grid, dim3(32, 8), 0, bp)(d.d_vol, d_zy, g->nx, g->ny, nz_i);
            launches++;
            if (cudaGetLastError() != cudaSuccess ||
                cudaMemcpy2DAsync(vol_zy + d.z_lo, (size_t)g->nz * sizeof(float), d_zy, (size_t)nz_i * sizeof(float), (size_t)nz_i * sizeof(float),
                                  slice, cudaMemcpyDeviceToHost, bp) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "transposed slab", __FILE__, __LINE__);
        }
    }
    // ---- the filtered maps, if wanted: every device sends its own chunks (dense copies made right after the filter)
    for (int i = 0; i < nd && rc == MONTE_OK && filtered; i++) {
        FdkMultiDev &d = dv[i];
        if ((rc = use_dev(i))) break;
        cudaStream_t st = ctx().stream;
        for (int k = 0; k < n_up && rc == MONTE_OK; k++) {
            const int ch = k * nd + i, a = V[ch], b = V[ch + 1];
            if (b <= a) continue;
            if (cudaMemcpyAsync(filtered + (size_t)a * per_view, d.d_map + (size_t)d.off[k] * per_view, (size_t)(b - a) * per_view * sizeof(float),
                                cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "download of the filtered views", __FILE__, __LINE__);
        }
    }
    for (int i = 0; i < nd; i++) {                                   // (also after an error: nothing may stay in flight)
        if (use_dev(i) != MONTE_OK) continue;
        if (rc != MONTE_OK && zs_pending[i]) bp_join_zstreams(ctx().aux_stream, zs_pending[i]);
        cudaError_t e = cudaStreamSynchronize(ctx().stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().aux_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().copy_stream);
        if (e != cudaSuccess && rc == MONTE_OK) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    if (rc == MONTE_OK && stats) {
        memset(stats, 0, sizeof(*stats));
        for (int i = 0; i < nd; i++) {
            use_dev(i);
            float tf = 0.f, tb = 0.f;
            // ms_filter: upload + filter of the device's chunks; ms_backproject: what the backprojection still needed after
            // that (the rest ran underneath the uploads); both as the maximum over the devices
            if (dv[i].start && cudaEventElapsedTime(&tf, dv[i].start, dv[i].filter_end) == cudaSuccess) stats->ms_filter = fmax(stats->ms_filter, (double)tf);
            if (dv[i].start && cudaEventElapsedTime(&tb, dv[i].start, dv[i].bp_end) == cudaSuccess) stats->ms_backproject = fmax(stats->ms_backproject, (double)tb);
        }
        cudaGetLastError();
        stats->ms_backproject = fmax(0.0, stats->ms_backproject - stats->ms_filter);
        stats->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        stats->ms_d2h = fmax(0.0, stats->ms_total - stats->ms_filter - stats->ms_backproject);   // what compute did not hide (host clock)
        stats->voxel_updates = (uint64_t)(g->s_end - g->s_begin) * (g->t_end - g->t_begin) * (g->z_end - g->z_begin) * g->n_views;
        stats->filter_macs = (uint64_t)g->n_views * g->nv * g->nu * g->nu;
        stats->launches = launches; stats->sm_count = ctx_of(0).sm_count;
    }
    for (int ch = 0; ch < C; ch++) if (ev_f[ch]) cudaEventDestroy(ev_f[ch]);
    use_dev(0);
    return rc;
}

// Host-buffer pipeline: the drop-in for bp3d20.cpp:29-171.
int monte_gpu_fdk(const monte_fdk_geom *g, const float *map, float *filtered, float *vol_xy, float *vol_zy,
                  monte_fdk_stats *stats) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_geom(g)) return rc;
    MONTE_ARG(map && vol_xy, "fdk: map and vol_xy must not be NULL");
    if (n_dev() > 1) return fdk_multi(g, map, filtered, vol_xy, vol_zy, stats);
    Context &c = ctx();
    cudaStream_t st = c.stream, cp = c.copy_stream;
    const size_t per_view = (size_t)g->nu * g->nv;
    const size_t n_map = (size_t)g->n_views * per_view;
    const size_t slice = (size_t)g->nx * g->ny, n_vol = slice * g->nz;
    float *d_map = (float *)scratch(0, n_map * sizeof(float));
    float *d_filt = (float *)scratch(1, monte_gpu_fdk_filtered_elems(g) * sizeof(float));
    float *d_vol = (float *)scratch(2, n_vol * sizeof(float));
    if (!d_map || !d_filt || !d_vol) return MONTE_E_NOMEM;
    if (int rc = fdk_prepare(g, st)) return rc;
    // Two streams: the copy stream uploads view chunks while the compute stream filters the previous
    // chunk, and downloads finished z-slabs while the next slab is backprojected.  With pinned host
    // buffers the copies are truly asynchronous; with pageable ones they still overlap the kernels.
    constexpr int MAXC = FDK_MAXC;
    FdkDevState &ds = g_fdk_state.get();
    cudaEvent_t *ev_up = ds.ev_up, *ev_slab = ds.ev_slab, *ev_t = ds.ev_t;   // destroyed by fdk_cleanup at shutdown
    if (!ev_up[0]) {
        for (int i = 0; i < MAXC; i++) {
            MONTE_CUDA(cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming));
            MONTE_CUDA(cudaEventCreateWithFlags(&ev_slab[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 6; i++) MONTE_CUDA(cudaEventCreate(&ev_t[i]));
    }
    int launches = 0;
    MONTE_CUDA(cudaEventRecord(ev_t[0], st));                       // start of everything
    MONTE_CUDA(cudaStreamWaitEvent(cp, ev_t[0], 0));
    // Views travel in chunks: upload k | filter k | (pad k-1, which needs the first row of chunk k) |
    // backproject k-1 into the whole volume, continuing the stored partial sums.  The last `m_tail` chunks are
    // backprojected slab by slab (slab-major: all of those chunks into slab 0, then slab 1, ...) so that a finished slab
    // goes home while the next one is computed: the tail chunks together must outlast the download of the volume.
    int n_up = g->n_views >= 128 ? 16 : (g->n_views >= 64 ? 8 : 1);
    { const char *e = getenv("MONTE_FDK_CHUNKS"); if (e && atoi(e) >= 1 && atoi(e) <= MAXC && atoi(e) <= g->n_views) n_up = atoi(e); }   // (A/B knob)
    const int rows = g->n_views * g->nv, pitch = (int)filtered_pitch(g);
    const int n_slab = g->nz >= 128 ? 4 : 1;
    int m_tail = n_slab > 1 ? (n_up + 4) / 5 : 1;                  // ~20 % of the views
    { const char *e = getenv("MONTE_FDK_TAIL"); if (e && atoi(e) >= 1) m_tail = atoi(e); }
    if (m_tail > n_up) m_tail = n_up;
    auto chunk_lo = [&](int k) { return (int)((long long)g->n_views * k / n_up); };
    auto pad_chunk = [&](int k) -> int {                           // chunk k = views [chunk_lo(k), chunk_lo(k + 1)); needs chunk k + 1 filtered
        // the last view of the chunk reads up to two rows into the next chunk (already filtered):
        // their duplicated column must be valid too
        const int r0 = chunk_lo(k) * g->nv, r1 = k == n_up - 1 ? rows + 2 : min(chunk_lo(k + 1) * g->nv + 2, rows + 2);
        fdk_pad_kernel MONTE_CFG(ceil_div(r1 - r0, 256), 256, 0, st)(d_filt, rows, g->nu, pitch, r0, r1);
        MONTE_CUDA(cudaGetLastError());
        launches++;
        return MONTE_OK;
    };
    for (int k = 0; k < n_up; k++) {
        const int v0 = chunk_lo(k), v1 = chunk_lo(k + 1);
        MONTE_CUDA(cudaMemcpyAsync(d_map + v0 * per_view, map + v0 * per_view, (size_t)(v1 - v0) * per_view * sizeof(float),
                                   cudaMemcpyHostToDevice, cp));
        MONTE_CUDA(cudaEventRecord(ev_up[k], cp));
        MONTE_CUDA(cudaStreamWaitEvent(st, ev_up[k], 0));
        if (int rc = monte_gpu_fdk_filter_dev(g, d_map, v0, v1, d_filt, st)) return rc;
        launches++;
        if (k > 0 && k - 1 < n_up - m_tail) {                       // chunk k-1 into the whole volume
            if (int rc = pad_chunk(k - 1)) return rc;
            if (int rc = backproject_views(g, d_filt, 0, g->nz, d_vol, st, chunk_lo(k - 1), chunk_lo(k), k - 1 > 0)) return rc;
            launches++;
        }
    }
    for (int k = n_up - m_tail; k < n_up; k++) if (int rc = pad_chunk(k)) return rc;
    MONTE_CUDA(cudaEventRecord(ev_t[1], st));                       // everything filtered
    for (int q = 0; q < n_slab; q++) {
        int z0 = (int)((long long)g->nz * q / n_slab), z1 = (int)((long long)g->nz * (q + 1) / n_slab);
        z0 = q == 0 ? 0 : (z0 / 32) * 32;                           // slabs on z-block boundaries
        z1 = q == n_slab - 1 ? g->nz : (z1 / 32) * 32;
        if (z1 <= z0) continue;
        for (int k = n_up - m_tail; k < n_up; k++) {
            if (int rc = backproject_views(g, d_filt, z0, z1, d_vol + z0 * slice, st, chunk_lo(k), chunk_lo(k + 1), k > 0)) return rc;
            launches++;
        }
        MONTE_CUDA(cudaEventRecord(ev_slab[q], st));
        MONTE_CUDA(cudaStreamWaitEvent(cp, ev_slab[q], 0));
        MONTE_CUDA(cudaMemcpyAsync(vol_xy + z0 * slice, d_vol + z0 * slice, (size_t)(z1 - z0) * slice * sizeof(float),
                                   cudaMemcpyDeviceToHost, cp));
    }
    MONTE_CUDA(cudaEventRecord(ev_t[2], st));                       // backprojected
    if (filtered) {   // d_map is free now: reuse it for the dense copy of the filtered projections
        if (int rc = monte_gpu_fdk_unpad_dev(g, d_filt, d_map, st)) return rc;
        launches += 1;
        MONTE_CUDA(cudaMemcpyAsync(filtered, d_map, n_map * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (vol_zy) {
        float *d_zy = (float *)scratch(3, n_vol * sizeof(float));
        if (!d_zy) return MONTE_E_NOMEM;
        if (int rc = monte_gpu_fdk_transpose_dev(g, d_vol, d_zy, st)) return rc;
        launches += 1;
        MONTE_CUDA(cudaMemcpyAsync(vol_zy, d_zy, n_vol * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    MONTE_CUDA(cudaEventRecord(ev_t[3], st));
    MONTE_CUDA(cudaEventRecord(ev_t[4], cp));                       // last slab downloaded
    MONTE_CUDA(cudaStreamWaitEvent(st, ev_t[4], 0));
    MONTE_CUDA(cudaEventRecord(ev_t[5], st));                       // end of everything
    MONTE_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_t[0], ev_t[1]); stats->ms_filter = ms;        // incl. the overlapped uploads
        cudaEventElapsedTime(&ms, ev_t[1], ev_t[2]); stats->ms_backproject = ms;
        cudaEventElapsedTime(&ms, ev_t[2], ev_t[3]); stats->ms_transpose = ms;
        cudaEventElapsedTime(&ms, ev_t[2], ev_t[5]); stats->ms_d2h = ms;           // download not hidden behind compute
        cudaEventElapsedTime(&ms, ev_t[0], ev_t[5]); stats->ms_total = ms;
        stats->ms_h2d = 0;                                                          // overlapped with the filter
        stats->voxel_updates = (uint64_t)(g->s_end - g->s_begin) * (g->t_end - g->t_begin) * (g->z_end - g->z_begin) * g->n_views;
        stats->filter_macs = (uint64_t)g->n_views * g->nv * g->nu * g->nu;
        stats->launches = launches; stats->sm_count = c.sm_count;
    }
    return MONTE_OK;
}

}  // extern "C"
