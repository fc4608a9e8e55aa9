// fbp2.cu — 2-D fan-beam FBP, the role of recon/fbp2.cpp:36-152.
// The shipped program looks the filtered sinogram up with a nearest-neighbour index
// (fbp2.cpp:126,147), which is discontinuous in the detector coordinate, so the whole path runs
// in double with the reference's expression order (no FMA contraction); it is ~24 M updates.
#include "common.cuh"
#include <cmath>

namespace monte {

struct Fbp2View { double c, s, cn, sn, px, py, tanb, inv; };   // per view, host-computed

// filtered[v][b] = sum_c pw[v][c]*scale*ramp[nu-1-b+c], double accumulator (fbp2.cpp:63-70)
__global__ void fbp2_filter_kernel(const float *sino, const double *wtab, const double *ramp, float *filt,
                                   int nu, int n_views, double scale) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (b >= nu || v >= n_views) return;
    double tmp = 0.;
    for (int c = 0; c < nu; c++) {
        const float pw = (float)__dmul_rn((double)sino[v * nu + c], wtab[c]);   // fbp2.cpp:38
        tmp = __dadd_rn(tmp, __dmul_rn(__dmul_rn((double)pw, scale), ramp[nu - 1 - b + c]));
    }
    filt[v * nu + b] = (float)tmp;
}

struct Fbp2Params {
    const float *filt; const Fbp2View *vc; float *img;
    int nu, n_views, view_first, nx, ny, s_begin, s_end, t_begin, t_end;
    double x0, y0, vox, dsd, half_u, inv_du, wd, beta_span, out_scale;
};

__global__ void fbp2_backproject_kernel(const Fbp2Params p) {
    const int s = p.s_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const int t = p.t_begin + blockIdx.y * blockDim.y + threadIdx.y;
    if (s >= p.s_end || t >= p.t_end) return;
    const double X = __dadd_rn(p.x0, __dmul_rn((double)s, p.vox));
    const double Y = __dsub_rn(p.y0, __dmul_rn((double)t, p.vox));
    float acc = 0.f;
    for (int v = p.view_first; v < p.n_views; v++) {
        const Fbp2View c = p.vc[v];
        const double pv0 = __dsub_rn(X, c.px), pv1 = __dsub_rn(Y, c.py);
        const double r0 = __dsub_rn(__dmul_rn(pv0, c.cn), __dmul_rn(pv1, c.sn));
        const double r1 = __dadd_rn(__dmul_rn(pv0, c.sn), __dmul_rn(pv1, c.cn));
        const double to_det = __ddiv_rn(p.dsd, r0);
        const double u = __dmul_rn(r1, to_det);
        if (fabs(u) > p.half_u) continue;                                   // fbp2.cpp:120
        const int index_y = (int)__dmul_rn(-p.inv_du, __dsub_rn(u, p.half_u));   // fbp2.cpp:126
        const double ts = __dsub_rn(__dmul_rn(X, c.c), __dmul_rn(Y, c.s));  // fbp2.cpp:132-133
        const double tt = __dadd_rn(__dmul_rn(X, c.s), __dmul_rn(Y, c.c));
        double d = __dmul_rn(fabs(__dadd_rn(__dmul_rn(-c.tanb, tt), ts)), 1.0);
        d = __ddiv_rn(d, c.inv);                                            // inv holds sqrt(1+tan^2)
        if (ts < 0) d = -d;
        const long long fi = (long long)v * p.nu + index_y;
        const float fv = (fi >= 0 && fi < (long long)p.nu * p.n_views) ? p.filt[fi] : 0.f;
        const double e = __dsub_rn(p.wd, d);
        double o = __ddiv_rn(__dmul_rn(p.wd, p.wd), __dmul_rn(e, e));
        o = __dmul_rn(o, (double)fv);
        o = __dmul_rn(o, p.beta_span);
        o = __dmul_rn(o, 2.0);
        o = __dmul_rn(o, M_PI);
        o = __ddiv_rn(o, 360.0);
        acc = (float)__dadd_rn((double)acc, __dmul_rn(o, p.out_scale));     // fbp2.cpp:148
    }
    p.img[(size_t)p.nx * t + s] = acc;
}

}  // namespace monte

using namespace monte;

extern "C" int monte_gpu_fbp2(const monte_fdk_geom *g, int view_first, const float *sino, float *filtered,
                              float *image, monte_fdk_stats *stats) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(g && sino && image, "fbp2: NULL argument");
    MONTE_ARG(g->n_views > 0 && g->nu > 0 && g->nx > 0 && g->ny > 0, "fbp2: bad sizes");
    MONTE_ARG(0 <= view_first && view_first <= g->n_views, "fbp2: bad view_first");
    MONTE_ARG(0 <= g->s_begin && g->s_end <= g->nx && 0 <= g->t_begin && g->t_end <= g->ny, "fbp2: bad ROI");
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const int nu = g->nu, nviews = g->n_views;
    std::vector<double> wtab(nu), ramp(2 * nu - 1, 0.0);
    const double wd = g->weight_dist;
    for (int zeta = 0; zeta < nu; zeta++) wtab[zeta] = wd / (sqrt(pow(wd, 2) + pow(-g->half_u + g->du * zeta, 2)));
    ramp[nu - 1] = 0.25;
    for (int n = 1; n < nu; n++)
        if (n % 2) ramp[nu - 1 + n] = ramp[nu - 1 - n] = -1. / pow(n * M_PI, 2);
    std::vector<Fbp2View> vc(nviews);
    for (int v = 0; v < nviews; v++) {
        double beta = g->angle0_deg + g->angle_step_deg * (double)v;
        float start_x = (float)(-g->dso), start_y = 0;
        Fbp2View &w = vc[v];
        w.c = cos(M_PI * beta / 180); w.s = sin(M_PI * beta / 180);
        w.cn = cos(-1 * M_PI * beta / 180); w.sn = sin(-1 * M_PI * beta / 180);
        w.px = start_x * w.c - start_y * w.s; w.py = start_x * w.s + start_y * w.c;
        w.tanb = tan(beta);
        w.inv = sqrt(1 + pow(tan(beta), 2));
    }
    const size_t n_s = (size_t)nu * nviews, n_img = (size_t)g->nx * g->ny;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_sino = take(n_s * 4), o_filt = take(n_s * 4), o_img = take(n_img * 4), o_w = take(nu * 8),
                 o_r = take(ramp.size() * 8), o_vc = take(vc.size() * sizeof(Fbp2View));
    char *base = (char *)scratch(4, off);
    if (!base) return MONTE_E_NOMEM;
    float *d_sino = (float *)(base + o_sino), *d_filt = (float *)(base + o_filt), *d_img = (float *)(base + o_img);
    double *d_w = (double *)(base + o_w), *d_r = (double *)(base + o_r);
    Fbp2View *d_vc = (Fbp2View *)(base + o_vc);
    EventTimer t_all(st);
    t_all.start();
    MONTE_CUDA(cudaMemcpyAsync(d_sino, sino, n_s * 4, cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemcpyAsync(d_w, wtab.data(), nu * 8, cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemcpyAsync(d_r, ramp.data(), ramp.size() * 8, cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemcpyAsync(d_vc, vc.data(), vc.size() * sizeof(Fbp2View), cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemsetAsync(d_img, 0, n_img * 4, st));
    fbp2_filter_kernel MONTE_CFG(dim3(ceil_div(nu, 64), nviews), 64, 0, st)(d_sino, d_w, d_r, d_filt, nu, nviews, g->filter_scale);
    MONTE_CUDA(cudaGetLastError());
    Fbp2Params p;
    p.filt = d_filt; p.vc = d_vc; p.img = d_img;
    p.nu = nu; p.n_views = nviews; p.view_first = view_first; p.nx = g->nx; p.ny = g->ny;
    p.s_begin = g->s_begin; p.s_end = g->s_end; p.t_begin = g->t_begin; p.t_end = g->t_end;
    p.x0 = g->x0; p.y0 = g->y0; p.vox = g->vox; p.dsd = g->dsd; p.half_u = g->half_u; p.inv_du = 1.0 / g->du;
    p.wd = wd; p.beta_span = (double)(float)g->angle_step_deg; p.out_scale = g->out_scale;
    if (g->s_end > g->s_begin && g->t_end > g->t_begin) {
        dim3 grid(ceil_div(g->s_end - g->s_begin, 32), ceil_div(g->t_end - g->t_begin, 4));
        fbp2_backproject_kernel MONTE_CFG(grid, dim3(32, 4), 0, st)(p);
        MONTE_CUDA(cudaGetLastError());
    }
    if (filtered) MONTE_CUDA(cudaMemcpyAsync(filtered, d_filt, n_s * 4, cudaMemcpyDeviceToHost, st));
    MONTE_CUDA(cudaMemcpyAsync(image, d_img, n_img * 4, cudaMemcpyDeviceToHost, st));
    t_all.stop();
    MONTE_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->ms_total = t_all.ms();
        stats->voxel_updates = (uint64_t)(g->s_end - g->s_begin) * (g->t_end - g->t_begin) * (nviews - view_first);
        stats->filter_macs = (uint64_t)nviews * nu * nu;
        stats->launches = 2; stats->sm_count = c.sm_count;
    }
    return MONTE_OK;
}
