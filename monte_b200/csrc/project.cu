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
    const uint8_t *labels;       // [z][y][x]
    const uint8_t *labels_t;     // [z][x][y]: the copy walked by views whose rays run mostly along x
    int nx, ny, nz;
    float pitch, inv_pitch;
    float org[3], clip_lo[3], clip_hi[3];
    float mu[256];               // linear attenuation per label at the requested energy
    int view_begin, n_views_run, det_ny, det_nx;
    float pixel, half, dso, dsd;
    double angle0, angle_step;
    float *map;                  // [n_views_total][ny][nx], written at absolute view index
    // MACRO kernels: macro-cells of 2^mshift voxels per side; mcell[(cz*mgy + cy)*mgx + cx] = the label all voxels of
    // the cell share, or 255 if they differ
    const uint8_t *mcell;
    int mshift, mgx, mgy;
    // leaping: mrad[cell] = n means the (2n-1)^3 block of macro-cells centred on the cell shares the cell's label
    // (0 for mixed cells); a ray inside the cell can advance (n-1) cell sides without leaving that block
    const uint8_t *mrad;
    float leap_unit;             // one cell side in cm, minus a safety margin
};

// one erosion pass of the block radius: a cell of radius `pass` whose 26 neighbours carry its label with radius >= pass
// has radius pass + 1 (cells outside the grid count as different)
__global__ void __launch_bounds__(128)
macro_radius_kernel(const uint8_t *__restrict__ mcell, const uint8_t *__restrict__ rin, uint8_t *__restrict__ rout,
                    int mgx, int mgy, int mgz, int pass) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= mgx * mgy * mgz) return;
    int r = pass == 0 ? (mcell[c] != 255 ? 1 : 0) : rin[c];
    if (pass > 0 && r == pass) {
        const int cx = c % mgx, cy = (c / mgx) % mgy, cz = c / (mgx * mgy);
        bool grow = cx > 0 && cy > 0 && cz > 0 && cx < mgx - 1 && cy < mgy - 1 && cz < mgz - 1;
        const int l = mcell[c];
        for (int dz = -1; dz <= 1 && grow; dz++)
            for (int dy = -1; dy <= 1 && grow; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int nb = ((cz + dz) * mgy + cy + dy) * mgx + cx + dx;
                    if (mcell[nb] != l || rin[nb] < pass) { grow = false; break; }
                }
        if (grow) r = pass + 1;
    }
    rout[c] = (uint8_t)r;
}
constexpr int PROJ_LEAP_PASSES = 24;     // block radius up to 25 cells

// label shared by all voxels of a macro-cell (from the padded [z][y][x] copy), 255 = mixed (or the label 255 itself)
__global__ void __launch_bounds__(128)
macro_cell_kernel(const uint8_t *__restrict__ lab, uint8_t *__restrict__ mcell, int nx, int ny, int nz, int mshift,
                  int mgx, int mgy, int mgz) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= mgx * mgy * mgz) return;
    const int cx = c % mgx, cy = (c / mgx) % mgy, cz = c / (mgx * mgy);
    const int C = 1 << mshift, px = nx + 2, py = ny + 2;
    const int x1 = min((cx + 1) * C, nx), y1 = min((cy + 1) * C, ny), z1 = min((cz + 1) * C, nz);
    int first = -1;
    bool mixed = false;
    for (int z = cz * C; z < z1 && !mixed; z++)
        for (int y = cy * C; y < y1 && !mixed; y++) {
            const uint8_t *row = lab + ((size_t)(z + 1) * py + (y + 1)) * px + 1;
            for (int x = cx * C; x < x1; x++) {
                const int l = row[x];
                if (first < 0) first = l;
                else if (l != first) { mixed = true; break; }
            }
        }
    mcell[c] = (mixed || first < 0 || first == 255) ? 255 : (uint8_t)first;
}

// Raw labels [z][y][x] -> two copies with a one-voxel guard ring of air (label 0): `out` [z][y][x]
// and `out_t` [z][x][y].  32x32 byte tiles go through shared memory for the transposed one.  The
// guard ring makes the walk below safe without per-step bounds tests: a plane crossing that rounding
// puts a hair before the exit face steps into air, never out of the allocation.
__global__ void __launch_bounds__(256)
labels_pad_transpose_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint8_t *__restrict__ out_t,
                            int nx, int ny) {
    __shared__ uint8_t tile[32][33];
    const int px = nx + 2, py = ny + 2;
    const size_t slice = (size_t)blockIdx.z * nx * ny;
    const size_t pslice = (size_t)(blockIdx.z + 1) * px * py;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int x = x0 + threadIdx.x, y = y0 + r;
        if (x < nx && y < ny) {
            const uint8_t v = in[slice + (size_t)y * nx + x];
            tile[r][threadIdx.x] = v;
            out[pslice + (size_t)(y + 1) * px + x + 1] = v;
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int y = y0 + threadIdx.x, x = x0 + r;
        if (x < nx && y < ny) out_t[pslice + (size_t)(x + 1) * py + y + 1] = tile[threadIdx.x][r];
    }
}

// One thread per ray.  A warp holds 32 transaxially adjacent pixels of one detector row: at any
// step its rays sit in neighbouring voxels ACROSS the dominant direction of travel, so with the
// label copy whose fastest axis is that transverse axis a warp-wide fetch touches one or two
// 128-byte lines instead of 32.
// MACRO = true (MONTE_PROJ_MACRO=n, macro-cells of 2^n voxels): whenever the walk enters a macro-cell whose voxels
// all carry one label, the whole cell is crossed in one segment -- the DDA state of the three axes is advanced by the
// number of voxel faces passed -- instead of voxel by voxel.  Same line integral (the segments still tile the ray),
// far fewer steps through homogeneous regions.
template <bool MACRO>
__global__ void __launch_bounds__(128)
project_primary_kernel(const __grid_constant__ ProjParams p) {
    __shared__ float s_mu[256];                               // per-label attenuation (a divergent index into
    for (int l = threadIdx.x; l < 256; l += 128) s_mu[l] = p.mu[l];   // the kernel-parameter bank would serialise)
    __syncthreads();
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);       // transaxial pixel: lanes of a warp
    const int j = blockIdx.y * 4 + (threadIdx.x >> 5);        // axial pixel (fastest in the map)
    const int view = p.view_begin + blockIdx.z;
    if (i >= p.det_ny || j >= p.det_nx) return;
    double sb_d, cb_d;
    sincospi((p.angle0 + p.angle_step * view) / 180.0, &sb_d, &cb_d);
    const float cb = (float)cb_d, sb = (float)sb_d;
    const bool along_x = fabsf(cb) >= fabsf(sb);              // uniform per block
    const uint8_t *__restrict__ lab_base = along_x ? p.labels_t : p.labels;
    const float yl = p.half - p.pixel * ((float)i + 0.5f), zl = p.half - p.pixel * ((float)j + 0.5f);
    const float rn = rsqrtf(p.dsd * p.dsd + yl * yl + zl * zl);
    const float d[3] = {(p.dsd * cb - yl * sb) * rn, (p.dsd * sb + yl * cb) * rn, zl * rn};
    const float src[3] = {-p.dso * cb, -p.dso * sb, 0.f};
    float t0 = 0.f, t1 = 1e30f;
#pragma unroll
    for (int a = 0; a < 3; a++) {                             // branch-free slab clip (inf bounds for d == 0)
        const float inv = 1.0f / d[a];
        const float ta = (p.clip_lo[a] - src[a]) * inv, tb = (p.clip_hi[a] - src[a]) * inv;
        t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
    }
    float acc = 0.f;
    if (t0 < t1) {
        // Amanatides-Woo with the ray parameter measured from the entry point (keeps fp32 resolution
        // ~1e-6 cm), a linear voxel index advanced by the stride of the crossed axis, and the next
        // label fetched while the current segment is accumulated.
        const float len = t1 - t0;
        float tn[3], dt[3];
        int idx[3], stp[3];
        const int dims[3] = {p.nx, p.ny, p.nz};
        const int px = p.nx + 2, py = p.ny + 2;
        const int strides[3] = {along_x ? py : 1, along_x ? 1 : px, px * py};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float e = fmaf(t0, d[a], src[a]);
            const float q = fmaf(1e-4f * p.pitch, d[a], e);
            idx[a] = min(max((int)floorf((q - p.org[a]) * p.inv_pitch), 0), dims[a] - 1);
            if (d[a] != 0.f) {
                const float edge = p.org[a] + (float)(idx[a] + (d[a] > 0.f ? 1 : 0)) * p.pitch;
                tn[a] = fmaxf((edge - e) / d[a], 0.f);       // >= 0, so the crossings below never decrease
                dt[a] = p.pitch / fabsf(d[a]);
            } else { tn[a] = 1e30f; dt[a] = 1e30f; }
            stp[a] = d[a] > 0.f ? strides[a] : -strides[a];
        }
        unsigned lin = (unsigned)((idx[2] + 1) * strides[2] + (idx[1] + 1) * strides[1] + (idx[0] + 1) * strides[0]);
        int lab = __ldg(lab_base + lin);
        float t = 0.f;
        bool check = true;                                    // MACRO: a macro-cell was entered since the last look-up
        const int cmask = (1 << p.mshift) - 1;
        while (t < len) {
            if (MACRO && check) {
                check = false;
                if ((unsigned)idx[0] < (unsigned)p.nx && (unsigned)idx[1] < (unsigned)p.ny && (unsigned)idx[2] < (unsigned)p.nz) {
                    const int cl = __ldg(p.mcell + ((idx[2] >> p.mshift) * p.mgy + (idx[1] >> p.mshift)) * p.mgx + (idx[0] >> p.mshift));
                    if (cl != 255) {
                        const int nr = __ldg(p.mrad + ((idx[2] >> p.mshift) * p.mgy + (idx[1] >> p.mshift)) * p.mgx + (idx[0] >> p.mshift));
                        if (nr >= 2) {
                            // deep inside a homogeneous block: leap (nr - 1) cell sides in one segment, then set the
                            // voxel walk up again at the landing point (as at the entry point above)
                            const float te = fminf(fmaf((float)(nr - 1), p.leap_unit, t), len);
                            acc = fmaf(s_mu[cl], te - t, acc);
                            t = te;
                            if (te >= len) break;
#pragma unroll
                            for (int a = 0; a < 3; a++) {
                                const float e = fmaf(t0 + t, d[a], src[a]);
                                const float q = fmaf(1e-4f * p.pitch, d[a], e);
                                idx[a] = min(max((int)floorf((q - p.org[a]) * p.inv_pitch), 0), dims[a] - 1);
                                if (d[a] != 0.f) {
                                    const float edge = p.org[a] + (float)(idx[a] + (d[a] > 0.f ? 1 : 0)) * p.pitch;
                                    tn[a] = t + fmaxf((edge - e) / d[a], 0.f);
                                }
                            }
                            lin = (unsigned)((idx[2] + 1) * strides[2] + (idx[1] + 1) * strides[1] + (idx[0] + 1) * strides[0]);
                            lab = __ldg(lab_base + lin);
                            check = true;
                            continue;
                        }
                        // faces left to the far side of the cell along each axis, and where the ray leaves the cell
                        int k[3];
                        float tx[3];
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            const int in_cell = idx[a] & cmask;
                            const int last = min(cmask, dims[a] - 1 - (idx[a] - in_cell));      // cells at the far edge are smaller
                            k[a] = d[a] > 0.f ? last - in_cell : in_cell;
                            tx[a] = fmaf(dt[a], (float)k[a], tn[a]);                            // 1e30 stays huge for d == 0
                        }
                        const float tex = fminf(tx[0], fminf(tx[1], tx[2]));
                        const float te = fminf(tex, len);
                        acc = fmaf(s_mu[cl], te - t, acc);
                        t = te;
                        if (tex >= len) break;
                        // advance every axis by the faces it passed up to tex (the exit axis: all its k + 1)
                        const int ea = tx[0] <= fminf(tx[1], tx[2]) ? 0 : (tx[1] <= tx[2] ? 1 : 2);
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            int n = 0;
                            if (a == ea) n = k[a] + 1;
                            else if (tn[a] <= tex) n = min((int)((tex - tn[a]) / dt[a]) + 1, k[a]);
                            tn[a] = fmaf(dt[a], (float)n, tn[a]);
                            idx[a] += d[a] > 0.f ? n : -n;
                            lin += (unsigned)(stp[a] * n);
                        }
                        lab = __ldg(lab_base + lin);
                        check = true;
                        continue;
                    }
                }
            }
            const float m12 = fminf(tn[1], tn[2]);
            const bool ax = tn[0] <= m12;
            const bool ay = !ax && tn[1] <= tn[2];
            const float te = fminf(fminf(tn[0], m12), len);
            acc = fmaf(s_mu[lab], te - t, acc);
            t = te;
            lin += (unsigned)(ax ? stp[0] : (ay ? stp[1] : stp[2]));
            lab = __ldg(lab_base + lin);                      // at most one voxel past the box: the guard ring
            const bool az = !ax && !ay;
            if (ax) tn[0] += dt[0];
            if (ay) tn[1] += dt[1];
            if (az) tn[2] += dt[2];
            if (MACRO) {                                      // keep the voxel index; a new macro-cell is worth a look-up
                if (ax) { idx[0] += d[0] > 0.f ? 1 : -1; check = (idx[0] & cmask) == (d[0] > 0.f ? 0 : cmask); }
                if (ay) { idx[1] += d[1] > 0.f ? 1 : -1; check = (idx[1] & cmask) == (d[1] > 0.f ? 0 : cmask); }
                if (az) { idx[2] += d[2] > 0.f ? 1 : -1; check = (idx[2] & cmask) == (d[2] > 0.f ? 0 : cmask); }
            }
        }
    }
    p.map[((size_t)view * p.det_ny + i) * p.det_nx + j] = acc;
}

}  // namespace monte

using namespace monte;

// block radius of every macro-cell by PROJ_LEAP_PASSES ping-pong erosion passes; *out = the buffer holding the result
static int macro_radius(const uint8_t *d_cell, uint8_t *a, uint8_t *b, int mgx, int mgy, int mgz, cudaStream_t st, const uint8_t **out) {
    const int n = mgx * mgy * mgz;
    uint8_t *src = b, *dst = a;
    for (int pass = 0; pass <= PROJ_LEAP_PASSES; pass++) {
        macro_radius_kernel MONTE_CFG(ceil_div(n, 128), 128, 0, st)(d_cell, src, dst, mgx, mgy, mgz, pass);
        MONTE_CUDA(cudaGetLastError());
        uint8_t *t = src; src = dst; dst = t;
    }
    *out = src;
    return MONTE_OK;
}

extern "C" int monte_gpu_project_primary(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                                         const monte_mc_xs *xs, double keV, int view_begin, int view_end, float *map) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(g && vol && labels && xs && map, "project_primary: NULL argument");
    MONTE_ARG(g->n_views > 0 && g->ny > 0 && g->nx > 0 && g->pixel > 0, "project_primary: bad detector");
    MONTE_ARG(vol->nx > 0 && vol->ny > 0 && vol->nz > 0 && vol->pitch > 0, "project_primary: bad volume");
    MONTE_ARG(xs->n_materials >= 1 && xs->n_materials <= MONTE_MC_MAX_MATERIALS, "project_primary: bad materials");
    if (view_end < 0) { MONTE_ARG(view_begin == 0, "project_primary: view_end < 0 (all views) needs view_begin == 0"); view_end = g->n_views; }
    MONTE_ARG(0 <= view_begin && view_begin <= view_end && view_end <= g->n_views, "project_primary: bad view range");
    if (view_begin == view_end) return MONTE_OK;
    cudaStream_t st = ctx().stream;
    const size_t nvox = (size_t)vol->nx * vol->ny * vol->nz;
    const size_t n_map = (size_t)g->n_views * g->ny * g->nx;
    const size_t nvox_al = (nvox + 255) / 256 * 256;
    const size_t npad = (size_t)(vol->nx + 2) * (vol->ny + 2) * (vol->nz + 2);
    const size_t npad_al = (npad + 255) / 256 * 256;
    MONTE_ARG(npad < ((size_t)1 << 32), "project_primary: volume too large for 32-bit voxel offsets");
    // tuning knob: MONTE_PROJ_MACRO=n (1..6) crosses homogeneous macro-cells of 2^n voxels in one segment (0 = off)
    static int macro = -1;
    if (macro < 0) { const char *e = getenv("MONTE_PROJ_MACRO"); macro = e ? atoi(e) : 0; if (macro < 0 || macro > 6) macro = 0; }
    const int mgx = macro ? ceil_div(vol->nx, 1 << macro) : 0, mgy = macro ? ceil_div(vol->ny, 1 << macro) : 0,
              mgz = macro ? ceil_div(vol->nz, 1 << macro) : 0;
    const size_t ncell_al = ((size_t)mgx * mgy * mgz + 255) / 256 * 256;
    char *base = (char *)scratch(6, nvox_al + 2 * npad_al + 3 * ncell_al + n_map * sizeof(float));
    if (!base) return MONTE_E_NOMEM;
    uint8_t *d_raw = (uint8_t *)base, *d_lab = d_raw + nvox_al, *d_lab_t = d_lab + npad_al, *d_cell = d_lab_t + npad_al;
    float *d_map = (float *)(base + nvox_al + 2 * npad_al + 3 * ncell_al);
    MONTE_CUDA(cudaMemcpyAsync(d_raw, labels, nvox, cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaMemsetAsync(d_lab, 0, 2 * npad_al, st));
    labels_pad_transpose_kernel MONTE_CFG(dim3(ceil_div(vol->nx, 32), ceil_div(vol->ny, 32), vol->nz), dim3(32, 8), 0, st)(
        d_raw, d_lab, d_lab_t, vol->nx, vol->ny);
    MONTE_CUDA(cudaGetLastError());
    const uint8_t *d_rad = nullptr;
    if (macro) {
        macro_cell_kernel MONTE_CFG(ceil_div(mgx * mgy * mgz, 128), 128, 0, st)(d_lab, d_cell, vol->nx, vol->ny, vol->nz, macro, mgx, mgy, mgz);
        MONTE_CUDA(cudaGetLastError());
        if (int rc = macro_radius(d_cell, d_cell + ncell_al, d_cell + 2 * ncell_al, mgx, mgy, mgz, st, &d_rad)) return rc;
    }
    ProjParams p;
    p.mcell = d_cell; p.mshift = macro; p.mgx = mgx; p.mgy = mgy;
    p.mrad = d_rad; p.leap_unit = (float)((double)(1 << macro) * vol->pitch * 0.999);
    p.labels = d_lab; p.labels_t = d_lab_t; p.nx = vol->nx; p.ny = vol->ny; p.nz = vol->nz;
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
    // views in chunks: chunk k is downloaded on the copy stream while chunk k+1 is traced
    constexpr int MAXC = 8;
    static cudaEvent_t ev[MAXC] = {nullptr};
    if (!ev[0]) {
        for (int i = 0; i < MAXC; i++) MONTE_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        at_shutdown([] { for (int i = 0; i < MAXC; i++) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; } });
    }
    cudaStream_t cp = ctx().copy_stream;
    const size_t per_view = (size_t)g->ny * g->nx;
    const int n_run = view_end - view_begin;
    const int chunk = ceil_div(n_run, MAXC);
    for (int k = 0, v0 = view_begin; v0 < view_end; k++, v0 += chunk) {
        const int v1 = v0 + chunk < view_end ? v0 + chunk : view_end;
        p.view_begin = v0; p.n_views_run = v1 - v0;
        dim3 grid(ceil_div(g->ny, 32), ceil_div(g->nx, 4), v1 - v0);
        if (macro) project_primary_kernel<true> MONTE_CFG(grid, 128, 0, st)(p);
        else project_primary_kernel<false> MONTE_CFG(grid, 128, 0, st)(p);
        MONTE_CUDA(cudaGetLastError());
        MONTE_CUDA(cudaEventRecord(ev[k], st));
        MONTE_CUDA(cudaStreamWaitEvent(cp, ev[k], 0));
        MONTE_CUDA(cudaMemcpyAsync(map + v0 * per_view, d_map + v0 * per_view, (size_t)(v1 - v0) * per_view * sizeof(float),
                                   cudaMemcpyDeviceToHost, cp));
    }
    MONTE_CUDA(cudaStreamSynchronize(st));
    MONTE_CUDA(cudaStreamSynchronize(cp));
    return MONTE_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-resident projector: the label copies are built once, line integrals land in a device buffer
// (the layout monte_gpu_fdk_filter_dev reads), so projection -> FDK runs without a host round trip and
// a multi-GPU host projects exactly the views it filters (no exchange before the filter).
// ------------------------------------------------------------------------------------------------
struct monte_projector {
    monte_mc_volume vol;
    uint8_t *d_raw = nullptr, *d_lab = nullptr, *d_lab_t = nullptr, *d_cell = nullptr;
    const uint8_t *d_rad = nullptr;
    int macro = 0, mgx = 0, mgy = 0, mgz = 0;
};

static int macro_knob() {
    const char *e = getenv("MONTE_PROJ_MACRO");
    const int m = e ? atoi(e) : 0;
    return m < 0 || m > 6 ? 0 : m;
}

extern "C" void monte_gpu_projector_destroy(monte_projector *s) {
    if (!s) return;
    cudaFree(s->d_raw);                        // one allocation holds all four buffers
    delete s;
}

extern "C" int monte_gpu_projector_create(const monte_mc_volume *vol, const uint8_t *labels, monte_projector **out) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(vol && labels && out, "projector_create: NULL argument");
    MONTE_ARG(vol->nx > 0 && vol->ny > 0 && vol->nz > 0 && vol->pitch > 0, "projector_create: bad volume");
    const size_t nvox = (size_t)vol->nx * vol->ny * vol->nz;
    const size_t nvox_al = (nvox + 255) / 256 * 256;
    const size_t npad = (size_t)(vol->nx + 2) * (vol->ny + 2) * (vol->nz + 2);
    const size_t npad_al = (npad + 255) / 256 * 256;
    MONTE_ARG(npad < ((size_t)1 << 32), "projector_create: volume too large for 32-bit voxel offsets");
    for (int a = 0; a < 3; a++) MONTE_ARG(vol->clip_lo[a] < vol->clip_hi[a], "projector_create: empty clip box");
    monte_projector *s = new monte_projector();
    s->vol = *vol;
    s->macro = macro_knob();
    if (s->macro) {
        s->mgx = ceil_div(vol->nx, 1 << s->macro); s->mgy = ceil_div(vol->ny, 1 << s->macro); s->mgz = ceil_div(vol->nz, 1 << s->macro);
    }
    const size_t ncell_al = ((size_t)s->mgx * s->mgy * s->mgz + 255) / 256 * 256;
    cudaStream_t st = ctx().stream;
    cudaError_t e = cudaMalloc(&s->d_raw, nvox_al + 2 * npad_al + 3 * ncell_al + 256);
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMalloc(projector)", __FILE__, __LINE__); }
    s->d_lab = s->d_raw + nvox_al; s->d_lab_t = s->d_lab + npad_al; s->d_cell = s->d_lab_t + npad_al;
    int rc = MONTE_OK;
    do {
        if ((e = cudaMemcpyAsync(s->d_raw, labels, nvox, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(s->d_lab, 0, 2 * npad_al, st)) != cudaSuccess) break;
        labels_pad_transpose_kernel MONTE_CFG(dim3(ceil_div(vol->nx, 32), ceil_div(vol->ny, 32), vol->nz), dim3(32, 8), 0, st)(
            s->d_raw, s->d_lab, s->d_lab_t, vol->nx, vol->ny);
        if (s->macro)
            macro_cell_kernel MONTE_CFG(ceil_div(s->mgx * s->mgy * s->mgz, 128), 128, 0, st)(s->d_lab, s->d_cell, vol->nx, vol->ny, vol->nz,
                                                                                       s->macro, s->mgx, s->mgy, s->mgz);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if (s->macro && macro_radius(s->d_cell, s->d_cell + ncell_al, s->d_cell + 2 * ncell_al, s->mgx, s->mgy, s->mgz, st, &s->d_rad)) {
            e = cudaErrorInvalidValue;             // (the message of the failed launch is already set)
            break;
        }
        e = cudaStreamSynchronize(st);             // `labels` may be pageable; the scene is ready when this returns
    } while (0);
    if (e != cudaSuccess) { rc = cuda_fail(e, "projector_create", __FILE__, __LINE__); monte_gpu_projector_destroy(s); return rc; }
    *out = s;
    return MONTE_OK;
}

extern "C" int monte_gpu_project_primary_dev(const monte_projector *s, const monte_mc_geom *g, const monte_mc_xs *xs, double keV,
                                             int view_begin, int view_end, float *d_map, void *stream) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(s && g && xs && d_map, "project_primary_dev: NULL argument");
    MONTE_ARG(g->n_views > 0 && g->ny > 0 && g->nx > 0 && g->pixel > 0, "project_primary_dev: bad detector");
    MONTE_ARG(xs->n_materials >= 1 && xs->n_materials <= MONTE_MC_MAX_MATERIALS, "project_primary_dev: bad materials");
    MONTE_ARG(0 <= view_begin && view_begin <= view_end && view_end <= g->n_views, "project_primary_dev: bad view range");
    if (view_begin == view_end) return MONTE_OK;
    const monte_mc_volume *vol = &s->vol;
    ProjParams p;
    p.labels = s->d_lab; p.labels_t = s->d_lab_t; p.nx = vol->nx; p.ny = vol->ny; p.nz = vol->nz;
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
    p.mcell = s->d_cell; p.mshift = s->macro; p.mgx = s->mgx; p.mgy = s->mgy;
    p.mrad = s->d_rad; p.leap_unit = (float)((double)(1 << s->macro) * vol->pitch * 0.999);
    cudaStream_t st = (cudaStream_t)stream;
    // grid.z carries the views: at most 65535 per launch
    for (int v0 = view_begin; v0 < view_end; v0 += 32768) {
        const int v1 = v0 + 32768 < view_end ? v0 + 32768 : view_end;
        p.view_begin = v0; p.n_views_run = v1 - v0;
        dim3 grid(ceil_div(g->ny, 32), ceil_div(g->nx, 4), v1 - v0);
        if (s->macro) project_primary_kernel<true> MONTE_CFG(grid, 128, 0, st)(p);
        else project_primary_kernel<false> MONTE_CFG(grid, 128, 0, st)(p);
        MONTE_CUDA(cudaGetLastError());
    }
    return MONTE_OK;
}
