// project.cu — deterministic cone-beam primary projection through the label volume.
//
// The reference only has a 2-D ray-march (monte_cpp/projection.cpp:74-125) and a 3-D draft that
// does not compile (monte_cpp/3d_projection.cpp); its MC primaries (image0 of
// CBCT_real325im.cu:584) estimate exp(-integral mu dl) along the source->pixel-centre ray.  This
// kernel computes that integral exactly (Amanatides-Woo voxel traversal), one thread per detector
// pixel and view: the variance-free input for FDK (BASELINE config 3) and the 1e-4 check of the
// MC primary tally.
#include "common.cuh"
#include <cmath>

namespace monte {

struct ProjParams {
    const uint8_t *labels;
    int nx, ny, nz;
    float pitch, inv_pitch;
    float org[3], clip_lo[3], clip_hi[3];
    float mu[256];               // linear attenuation per label at the requested energy
    int view_begin, n_views_run, det_ny, det_nx;
    float pixel, half, dso, dsd;
    double angle0, angle_step;
    float *map;                  // [n_views_total][ny][nx], written at absolute view index
};

__global__ void __launch_bounds__(128)
project_primary_kernel(const __grid_constant__ ProjParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;      // axial pixel (fastest)
    const int i = blockIdx.y;
    const int view = p.view_begin + blockIdx.z;
    if (j >= p.det_nx) return;
    double sb_d, cb_d;
    sincospi((p.angle0 + p.angle_step * view) / 180.0, &sb_d, &cb_d);
    const float cb = (float)cb_d, sb = (float)sb_d;
    const float yl = p.half - p.pixel * ((float)i + 0.5f), zl = p.half - p.pixel * ((float)j + 0.5f);
    const float rn = rsqrtf(p.dsd * p.dsd + yl * yl + zl * zl);
    const float d[3] = {(p.dsd * cb - yl * sb) * rn, (p.dsd * sb + yl * cb) * rn, zl * rn};
    const float src[3] = {-p.dso * cb, -p.dso * sb, 0.f};
    float t0 = 0.f, t1 = 1e30f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (d[a] != 0.f) {
            const float inv = 1.0f / d[a];
            float ta = (p.clip_lo[a] - src[a]) * inv, tb = (p.clip_hi[a] - src[a]) * inv;
            if (ta > tb) { const float t = ta; ta = tb; tb = t; }
            t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
        } else if (src[a] < p.clip_lo[a] || src[a] >= p.clip_hi[a]) t1 = -1.f;
    }
    float acc = 0.f;
    if (t0 < t1) {
        // traverse with the parameter measured from the entry point (keeps fp32 resolution ~1e-6 cm)
        const float len = t1 - t0;
        float e[3], tnext[3], dt[3];
        int idx[3], step[3];
        const int dims[3] = {p.nx, p.ny, p.nz};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            e[a] = fmaf(t0, d[a], src[a]);
            const float q = fmaf(1e-4f * p.pitch, d[a], e[a]);
            idx[a] = (int)floorf((q - p.org[a]) * p.inv_pitch);
            step[a] = d[a] > 0.f ? 1 : -1;
            if (d[a] != 0.f) {
                const float edge = p.org[a] + (float)(idx[a] + (d[a] > 0.f ? 1 : 0)) * p.pitch;
                tnext[a] = (edge - e[a]) / d[a];
                dt[a] = p.pitch / fabsf(d[a]);
            } else { tnext[a] = 1e30f; dt[a] = 1e30f; }
        }
        float t = 0.f;
        while (t < len) {
            const int a = tnext[0] <= tnext[1] ? (tnext[0] <= tnext[2] ? 0 : 2) : (tnext[1] <= tnext[2] ? 1 : 2);
            const float te = fminf(tnext[a], len);
            if (idx[0] >= 0 && idx[1] >= 0 && idx[2] >= 0 && idx[0] < dims[0] && idx[1] < dims[1] && idx[2] < dims[2]) {
                const int l = __ldg(p.labels + ((size_t)idx[2] * p.ny + idx[1]) * p.nx + idx[0]);
                if (te > t) acc = fmaf(p.mu[l], te - t, acc);
            }
            t = te;
            if (a == 0) { idx[0] += step[0]; tnext[0] += dt[0]; }
            else if (a == 1) { idx[1] += step[1]; tnext[1] += dt[1]; }
            else { idx[2] += step[2]; tnext[2] += dt[2]; }
        }
    }
    p.map[((size_t)view * p.det_ny + i) * p.det_nx + j] = acc;
}

}  // namespace monte

using namespace monte;

extern "C" int monte_gpu_project_primary(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                                         const monte_mc_xs *xs, double keV, int view_begin, int view_end, float *map) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(g && vol && labels && xs && map, "project_primary: NULL argument");
    MONTE_ARG(g->n_views > 0 && g->ny > 0 && g->nx > 0 && g->pixel > 0, "project_primary: bad detector");
    MONTE_ARG(vol->nx > 0 && vol->ny > 0 && vol->nz > 0 && vol->pitch > 0, "project_primary: bad volume");
    MONTE_ARG(xs->n_materials >= 1 && xs->n_materials <= MONTE_MC_MAX_MATERIALS, "project_primary: bad materials");
    if (view_begin == 0 && view_end == 0) view_end = g->n_views;
    MONTE_ARG(0 <= view_begin && view_begin <= view_end && view_end <= g->n_views, "project_primary: bad view range");
    if (view_begin == view_end) return MONTE_OK;
    cudaStream_t st = ctx().stream;
    const size_t nvox = (size_t)vol->nx * vol->ny * vol->nz;
    const size_t n_map = (size_t)g->n_views * g->ny * g->nx;
    char *base = (char *)scratch(6, nvox + 256 + n_map * sizeof(float));
    if (!base) return MONTE_E_NOMEM;
    uint8_t *d_lab = (uint8_t *)base;
    float *d_map = (float *)(base + (nvox + 255) / 256 * 256);
    MONTE_CUDA(cudaMemcpyAsync(d_lab, labels, nvox, cudaMemcpyHostToDevice, st));
    ProjParams p;
    p.labels = d_lab; p.nx = vol->nx; p.ny = vol->ny; p.nz = vol->nz;
    p.pitch = (float)vol->pitch; p.inv_pitch = (float)(1.0 / vol->pitch);
    for (int a = 0; a < 3; a++) { p.org[a] = (float)vol->origin[a]; p.clip_lo[a] = (float)vol->clip_lo[a]; p.clip_hi[a] = (float)vol->clip_hi[a]; }
    int k = (int)(keV + 0.5);
    k = k < 0 ? 0 : (k > MONTE_MC_TABLE_ROWS - 1 ? MONTE_MC_TABLE_ROWS - 1 : k);
    for (int l = 0; l < 256; l++) {
        const int m = l == 0 ? -1 : (l <= xs->n_materials ? l - 1 : xs->n_materials - 1);
        p.mu[l] = m < 0 ? 0.f : (float)((double)xs->total[m][k] * (double)xs->density[m]);
    }
    p.view_begin = view_begin; p.n_views_run = view_end - view_begin; p.det_ny = g->ny; p.det_nx = g->nx;
    p.pixel = (float)g->pixel; p.half = (float)g->half; p.dso = (float)g->dso; p.dsd = (float)(g->dso + g->dod);
    p.angle0 = g->angle0_deg; p.angle_step = g->angle_step_deg;
    p.map = d_map;
    dim3 grid(ceil_div(g->nx, 128), g->ny, view_end - view_begin);
    project_primary_kernel<<<grid, 128, 0, st>>>(p);
    MONTE_CUDA(cudaGetLastError());
    const size_t per_view = (size_t)g->ny * g->nx;
    MONTE_CUDA(cudaMemcpyAsync(map + view_begin * per_view, d_map + view_begin * per_view,
                               (size_t)(view_end - view_begin) * per_view * sizeof(float), cudaMemcpyDeviceToHost, st));
    MONTE_CUDA(cudaStreamSynchronize(st));
    return MONTE_OK;
}
