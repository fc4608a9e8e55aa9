// host_helpers.cpp — GPU-free pieces of the reference's driver surface: cross-section tables,
// phantom generators, CT-number conversion, geometry presets.  Linked into libmonte_gpu and used
// by the C++ drivers under monte_b200/host/.
#include <vector>
#include <thread>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "../../include/monte_gpu.h"

namespace monte { void set_error(const char *fmt, ...); }

extern "C" {

// ---- presets: the literals of the shipped recon programs ------------------------------------
static void fdk_common(monte_fdk_geom *g) {
    memset(g, 0, sizeof(*g));
    g->n_views = 360;                          // bp3d20.cpp:19
    g->half_u = g->half_v = 16.25;             // bp3d20.cpp:40,116
    g->dso = 160; g->dsd = 220;                // bp3d20.cpp:85,107
    g->weight_dist = 60;                       // bp3d20.cpp:40,159
    g->filter_scale = 0.5;                     // bp3d20.cpp:68
    g->out_scale = 2.7; g->out_scale2 = 1;     // bp3d20.cpp:160
    g->angle0_deg = 0; g->angle_step_deg = 1;  // bp3d20.cpp:82-83
    g->nx = g->ny = g->nz = 256; g->vox = 0.1; // bp3d20.cpp:18,96
    g->x0 = -12.8; g->y0 = 12.8; g->z0 = 12.8; // bp3d20.cpp:96-98
    g->s_begin = 125; g->s_end = 130;          // bp3d20.cpp:93
    g->t_begin = 0; g->t_end = 256; g->z_begin = 0; g->z_end = 256;
    g->mask_r2 = -1;
    g->weight_mode = MONTE_FDK_REFERENCE;
}
void monte_fdk_geom_bp3d20(monte_fdk_geom *g) {
    fdk_common(g);
    g->nu = g->nv = 65; g->du = g->dv = 0.5;               // bp3d20.cpp:17,40
    g->mask_cs = g->mask_ct = g->mask_cz = 128; g->mask_r2 = 118 * 118;  // bp3d20.cpp:145
    g->coord_mode = MONTE_FDK_COORD_SCALE_AFTER;
}
void monte_fdk_geom_bp3d20_325(monte_fdk_geom *g) {
    fdk_common(g);
    g->nu = g->nv = 325; g->du = g->dv = 0.1;              // bp3d20_325.cpp:17,43
    g->out_scale2 = 5;                                     // bp3d20_325.cpp:170
    g->coord_mode = MONTE_FDK_COORD_SCALE_BEFORE;          // bp3d20_325.cpp:134-135
}
void monte_fdk_geom_fbp2(monte_fdk_geom *g) {
    fdk_common(g);
    g->nu = 65; g->nv = 1; g->du = g->dv = 0.5;            // fbp2.cpp:17,38
    g->nz = 1; g->z_end = 1; g->s_begin = 0; g->s_end = 256;
    g->out_scale = 1.7;                                    // fbp2.cpp:148
}

// ---- cross-section tables ---------------------------------------------------------------------
// readcsv role (CBCT_real2.cpp:633-668): rows are "coh,compton,photo,total" for 1..200 keV.
int monte_xs_load_csv(const char *path, int material, float density, int quirk_bom, monte_mc_xs *xs) {
    if (!path || !xs || material < 0 || material >= MONTE_MC_MAX_MATERIALS) {
        monte::set_error("monte_xs_load_csv: bad argument");
        return MONTE_E_ARG;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        monte::set_error("monte_xs_load_csv: cannot open %s", path);
        return MONTE_E_IO;
    }
    char line[512];
    int row = 0;
    while (row < MONTE_MC_TABLE_ROWS - 1 && fgets(line, sizeof(line), f)) {
        char *p = line;
        if (row == 0 && (unsigned char)p[0] == 0xEF && (unsigned char)p[1] == 0xBB && (unsigned char)p[2] == 0xBF) p += 3;
        double a, b, c, d;
        if (sscanf(p, "%lf,%lf,%lf,%lf", &a, &b, &c, &d) != 4) {
            if (p[0] == '\r' || p[0] == '\n' || p[0] == 0) continue;
            fclose(f);
            monte::set_error("monte_xs_load_csv: %s line %d is not 4 comma-separated numbers", path, row + 1);
            return MONTE_E_IO;
        }
        const int k = row + 1;            // index = keV (CBCT_real2.cpp:646)
        xs->coh[material][k] = (float)a;
        xs->compt[material][k] = (float)b;
        xs->photo[material][k] = (float)c;
        xs->total[material][k] = (float)d;
        row++;
    }
    fclose(f);
    if (row != MONTE_MC_TABLE_ROWS - 1) {
        monte::set_error("monte_xs_load_csv: %s has %d rows, expected %d", path, row, MONTE_MC_TABLE_ROWS - 1);
        return MONTE_E_IO;
    }
    if (quirk_bom) xs->coh[material][1] = 1.372f;   // CBCT_real2.cpp:663, applied to every file
    xs->coh[material][0] = xs->coh[material][1];
    xs->compt[material][0] = xs->compt[material][1];
    xs->photo[material][0] = xs->photo[material][1];
    xs->total[material][0] = xs->total[material][1];
    xs->density[material] = density;
    if (xs->n_materials < material + 1) xs->n_materials = material + 1;
    return MONTE_OK;
}

// ---- phantoms ------------------------------------------------------------------------------------
void monte_make_fantom(uint8_t *g, int n, int cy, int cx, int r2) {   // make_fantom.cpp:10-19
    for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++)
            g[j * n + k] = ((j - cy) * (j - cy) + (k - cx) * (k - cx) <= r2) ? 1 : 0;
}

void monte_make_sphere(uint8_t *g, int nx, int ny, int nz, int cx, int cy, int cz, int r2) {  // make_image01.cpp:15-23
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++)
                g[((size_t)k * ny + j) * nx + i] =
                    ((i - cx) * (i - cx) + (j - cy) * (j - cy) + (k - cz) * (k - cz) <= r2) ? 1 : 0;
}

// ---- CT number -> mu -----------------------------------------------------------------------------
// The reference file only gets as far as mu_H2O = csv_H2O[3][(int)(E+0.5)]*dens (ctnum_to_mu.cpp:55);
// the conversion it is named after is the standard mu = mu_water*(1+HU/1000).
int monte_ctnum_to_mu(const float *hu, size_t n, const monte_mc_xs *xs, double keV,
                      float hu_air_max, float hu_bone_min, float *mu, uint8_t *labels) {
    if (!hu || !xs || xs->n_materials < 1) {
        monte::set_error("monte_ctnum_to_mu: bad argument");
        return MONTE_E_ARG;
    }
    int k = (int)(keV + 0.5);
    if (k < 1) k = 1;
    if (k > MONTE_MC_TABLE_ROWS - 1) k = MONTE_MC_TABLE_ROWS - 1;
    const float mu_w = xs->total[0][k] * xs->density[0];
    for (size_t i = 0; i < n; i++) {
        if (mu) {
            float m = mu_w * (1.0f + hu[i] / 1000.0f);
            mu[i] = m > 0.f ? m : 0.f;
        }
        if (labels) labels[i] = hu[i] <= hu_air_max ? 0 : (hu[i] >= hu_bone_min && xs->n_materials > 1 ? 2 : 1);
    }
    return MONTE_OK;
}

/* ---- N-class segmentation of a CT volume ----------------------------------------------------
 * The reference's ctnum_to_mu.cpp loads the water table, forms mu_H2O (:55) and stops; its transport
 * (CBCT_real325im.cu:640-646) reads a LABEL volume and picks Ca / H2O / PMMA tables by label.  This is the missing
 * link: HU -> labels the transport reads + the tables those labels index.                                      */
int monte_hu_classes_default(int have_calcium, monte_hu_class *c) {
    if (!c) { monte::set_error("monte_hu_classes_default: NULL argument"); return MONTE_E_ARG; }
    // hu_min, a, b, calcium mass fraction, density.  Densities follow rho ~ 1 + HU/1000 for soft tissue (the bin's
    // typical HU) and the ICRU-44 bone series above it (spongiosa 1.18 ... cortical 1.92 g/cm^3, Ca 22.5 % by mass).
    const monte_hu_class t[9] = {
        {-900.f, 0, -1, 0.f, 0.30f},      // lung
        {-400.f, 0, -1, 0.f, 0.93f},      // adipose
        {-40.f, 0, -1, 0.f, 1.01f},       // water / soft tissue
        {60.f, 0, -1, 0.f, 1.07f},        // muscle, organs with contrast
        {150.f, 0, 1, 0.050f, 1.18f},     // spongy bone
        {400.f, 0, 1, 0.120f, 1.40f},     // bone
        {800.f, 0, 1, 0.180f, 1.65f},     // dense bone
        {1300.f, 0, 1, 0.225f, 1.92f},    // cortical bone
        {0.f, 0, -1, 0.f, 0.f}};
    for (int i = 0; i < 8; i++) {
        c[i] = t[i];
        if (!have_calcium) { c[i].material_b = -1; c[i].frac_b = 0.f; }
    }
    return 8;
}

int monte_ctnum_segment(const float *hu, size_t n, const monte_hu_class *classes, int n_classes,
                        const monte_mc_xs *base, double keV, monte_mc_xs *out, uint8_t *labels, float *mu,
                        uint32_t *present) {
    if (!hu || !classes || !base || !out || !labels || n_classes < 1 || n_classes > 64 || base->n_materials < 1 ||
        base->n_materials > MONTE_MC_MAX_MATERIALS) {
        monte::set_error("monte_ctnum_segment: bad argument");
        return MONTE_E_ARG;
    }
    // classes -> materials
    int label_of[64];
    memset(out, 0, sizeof(*out));
    int nm = 0;
    for (int c = 0; c < n_classes; c++) {
        const monte_hu_class &k = classes[c];
        if (c > 0 && !(k.hu_min > classes[c - 1].hu_min)) { monte::set_error("monte_ctnum_segment: classes must ascend in hu_min (class %d)", c); return MONTE_E_ARG; }
        if (!(k.density > 0.f)) { label_of[c] = 0; continue; }
        if (k.material_a < 0 || k.material_a >= base->n_materials || k.material_b >= base->n_materials ||
            !(k.frac_b >= 0.f && k.frac_b <= 1.f)) { monte::set_error("monte_ctnum_segment: class %d names a material outside the base tables", c); return MONTE_E_ARG; }
        if (nm == MONTE_MC_MAX_MATERIALS) { monte::set_error("monte_ctnum_segment: more than %d non-air classes", MONTE_MC_MAX_MATERIALS); return MONTE_E_ARG; }
        const int a = k.material_a, b = k.material_b;
        const double fb = b >= 0 ? (double)k.frac_b : 0.0, fa = 1.0 - fb;
        for (int r = 0; r < MONTE_MC_TABLE_ROWS; r++) {       // mixture rule, per interaction type (mass coefficients)
            out->coh[nm][r] = (float)(fa * base->coh[a][r] + (b >= 0 ? fb * base->coh[b][r] : 0.0));
            out->compt[nm][r] = (float)(fa * base->compt[a][r] + (b >= 0 ? fb * base->compt[b][r] : 0.0));
            out->photo[nm][r] = (float)(fa * base->photo[a][r] + (b >= 0 ? fb * base->photo[b][r] : 0.0));
            out->total[nm][r] = (float)(fa * base->total[a][r] + (b >= 0 ? fb * base->total[b][r] : 0.0));
        }
        out->density[nm] = k.density;
        if (base->ff_points > 0) {                              // form factor of the main component (a stand-in anyway)
            memcpy(out->ff_x2[nm], base->ff_x2[a], sizeof(out->ff_x2[nm]));
            memcpy(out->ff_cum[nm], base->ff_cum[a], sizeof(out->ff_cum[nm]));
        }
        label_of[c] = ++nm;
    }
    if (nm == 0) { monte::set_error("monte_ctnum_segment: every class is air"); return MONTE_E_ARG; }
    out->n_materials = nm;
    out->ff_points = base->ff_points;
    int kr = (int)(keV + 0.5);
    kr = kr < 1 ? 1 : (kr > MONTE_MC_TABLE_ROWS - 1 ? MONTE_MC_TABLE_ROWS - 1 : kr);
    float mu_of[MONTE_MC_MAX_MATERIALS + 1] = {0.f};
    for (int m = 0; m < nm; m++) mu_of[m + 1] = out->total[m][kr] * out->density[m];
    // voxels: binary search over the ascending class edges, eight host threads for big volumes
    const int T = n >= (1u << 22) ? 8 : 1;
    uint32_t seen[8] = {0};
    auto work = [&](int t) {
        const size_t lo = n * t / T, hi = n * (t + 1) / T;
        uint32_t sm = 0;
        for (size_t i = lo; i < hi; i++) {
            const float h = hu[i];
            int a = -1, b = n_classes - 1;                      // largest c with hu_min[c] <= h, or -1 (NaN: -1 = air)
            if (!(h >= classes[0].hu_min)) b = -1;
            else { a = 0; while (a < b) { const int mid = (a + b + 1) >> 1; if (classes[mid].hu_min <= h) a = mid; else b = mid - 1; } }
            const int lab = b < 0 ? 0 : label_of[a];
            labels[i] = (uint8_t)lab;
            if (mu) mu[i] = mu_of[lab];
            if (lab) sm |= 1u << (lab - 1);
        }
        seen[t] = sm;
    };
    if (T == 1) work(0);
    else { std::thread th[8]; for (int t = 0; t < T; t++) th[t] = std::thread(work, t); for (int t = 0; t < T; t++) th[t].join(); }
    if (present) { *present = 0; for (int t = 0; t < T; t++) *present |= seen[t]; }
    return MONTE_OK;
}

/* Woodcock majorant per keV: max over the materials (labels 1..n_materials) that occur in `labels`
 * (all materials if labels is NULL) of total[m][k]*density[m] -- CBCT_real325im.cu:866 takes the max over
 * every table it loaded; restricting it to the materials present is what makes a calcium-free volume
 * cheap to track.  mu_max[k], k = 0..200, in 1/cm.                                                     */
int monte_xs_majorant(const monte_mc_xs *xs, const uint8_t *labels, size_t n, float *mu_max) {
    if (!xs || !mu_max || xs->n_materials < 1 || xs->n_materials > MONTE_MC_MAX_MATERIALS) {
        monte::set_error("monte_xs_majorant: bad argument");
        return MONTE_E_ARG;
    }
    bool present[MONTE_MC_MAX_MATERIALS];
    for (int m = 0; m < MONTE_MC_MAX_MATERIALS; m++) present[m] = labels == nullptr;
    for (size_t i = 0; labels && i < n; i++) {
        int l = labels[i];
        if (l == 0) continue;
        if (l > xs->n_materials) l = xs->n_materials;          // same clamp as the transport kernel
        present[l - 1] = true;
    }
    for (int k = 0; k < MONTE_MC_TABLE_ROWS; k++) {
        float mx = 0.f;
        for (int m = 0; m < xs->n_materials; m++)
            if (present[m]) { const float v = xs->total[m][k] * xs->density[m]; if (v > mx) mx = v; }
        mu_max[k] = mx;
    }
    return MONTE_OK;
}

/* Analytic stand-in for a measured Rayleigh form factor (the reference ships none, and has no deflection at
 * all: CBCT_real325im.cu:656-695): F(x)^2 ~ (1 + x^2/x0^2)^-4, the hydrogen-like 1s charge cloud, so that
 * integral_0^{x2} F^2 d(x^2) = x0^2/3 * (1 - (1 + x2/x0^2)^-3).  Grid: x2[0] = 0, then logarithmic from 1e-4
 * to 270 1/A^2 ((200 keV / 12.398 keV A)^2 = 260 is the largest x^2 a table row can ask for).               */
int monte_xs_formfactor_hydrogenic(monte_mc_xs *xs, int material, double x0) {
    if (!xs || material < 0 || material >= MONTE_MC_MAX_MATERIALS || !(x0 > 0)) {
        monte::set_error("monte_xs_formfactor_hydrogenic: bad argument");
        return MONTE_E_ARG;
    }
    const int n = MONTE_MC_FF_POINTS;
    const double lo = 1e-4, hi = 270.0, ratio = pow(hi / lo, 1.0 / (n - 2));
    for (int i = 0; i < n; i++) {
        const double x2 = i == 0 ? 0.0 : (i == n - 1 ? hi : lo * pow(ratio, i - 1));
        xs->ff_x2[material][i] = (float)x2;
        xs->ff_cum[material][i] = (float)(x0 * x0 / 3.0 * (1.0 - pow(1.0 + (double)(float)x2 / (x0 * x0), -3.0)));
    }
    xs->ff_points = n;
    return MONTE_OK;
}

/* ---- clearance grid for the two-level Woodcock majorant (monte_mc_volume.tracking_mode = CLEARANCE) -------
 * The reference tracks with one majorant, the maximum over every table it loaded (CBCT_real325im.cu:866-868).  At
 * diagnostic energies a dense insert (calcium) sets that maximum far above the attenuation of the water that
 * fills most of the volume, and most tentative collisions are virtual.  The clearance grid lets the tracker use
 * the majorant of the OTHER materials wherever the dense one is provably out of reach: cells of 2^cell_log2
 * voxels per side; grid[cell] = floor(2 * d) clipped to 127 (7 bits of slot state in the transport kernel), d = smallest distance, in cell sides, between the
 * cell's box and the box of any cell that contains a voxel of the heavy material (0 for those cells and their
 * neighbours).  A flight of at most grid[cell] * half a cell side that starts anywhere in the cell cannot reach
 * the heavy material.  Exact separable transform (the squared box distance is a sum over the axes).          */
int monte_mc_clearance_dims(const monte_mc_volume *vol, int cell_log2, int32_t dims[3]) {
    if (!vol || !dims || cell_log2 < 0 || cell_log2 > 8) { monte::set_error("monte_mc_clearance_dims: bad argument"); return MONTE_E_ARG; }
    const int c = 1 << cell_log2;
    dims[0] = (vol->nx + c - 1) >> cell_log2; dims[1] = (vol->ny + c - 1) >> cell_log2; dims[2] = (vol->nz + c - 1) >> cell_log2;
    return MONTE_OK;
}

// cells that hold the heavy material: 0, others INF
static void clearance_seed(const monte_mc_volume *vol, const uint8_t *labels, int n_materials, int heavy_material, int cell_log2,
                           int gx, int gy, std::vector<float> &f) {
    const float INF = 1e30f;
    std::fill(f.begin(), f.end(), INF);
    for (int z = 0; z < vol->nz; z++)
        for (int y = 0; y < vol->ny; y++) {
            const uint8_t *row = labels + ((size_t)z * vol->ny + y) * vol->nx;
            float *frow = f.data() + ((size_t)(z >> cell_log2) * gy + (y >> cell_log2)) * gx;
            for (int x = 0; x < vol->nx; x++) {
                int l = row[x];
                if (l == 0) continue;
                if (l > n_materials) l = n_materials;              // same clamp as the transport kernel
                if (l - 1 == heavy_material) frow[x >> cell_log2] = 0.f;
            }
        }
}

// f <- min over the cells j of the line of f[j] + max(0, |i - j| - 1)^2, one axis after the other; sign[a] = 0: every j,
// +1: only j >= i, -1: only j <= i (the cells a ray travelling in that direction along the axis can still reach)
static void clearance_transform(std::vector<float> &f, int gx, int gy, int gz, const int sign[3], uint8_t *grid) {
    const float INF = 1e30f;
    const size_t ncell = (size_t)gx * gy * gz;
    std::vector<float> t(ncell);
    const int dims[3] = {gx, gy, gz};
    const size_t stride[3] = {1, (size_t)gx, (size_t)gx * gy};
    for (int a = 0; a < 3; a++) {
        const int n = dims[a];
        const int b = (a + 1) % 3, c = (a + 2) % 3;
        for (int jb = 0; jb < dims[b]; jb++)
            for (int jc = 0; jc < dims[c]; jc++) {
                const size_t base = jb * stride[b] + jc * stride[c];
                for (int i = 0; i < n; i++) {
                    float best = f[base + i * stride[a]];
                    for (int k = 1; k < n; k++) {                      // outwards from i; farther cells cannot beat `best`
                        const float gap2 = (float)(k - 1) * (float)(k - 1);
                        if (gap2 >= best) break;
                        if (sign[a] <= 0 && i - k >= 0) { const float w = f[base + (i - k) * stride[a]] + gap2; if (w < best) best = w; }
                        if (sign[a] >= 0 && i + k < n) { const float w = f[base + (i + k) * stride[a]] + gap2; if (w < best) best = w; }
                    }
                    t[base + i * stride[a]] = best;
                }
            }
        f.swap(t);
    }
    for (size_t i = 0; i < ncell; i++) {
        const double q = f[i] >= INF ? 127.0 : floor(2.0 * sqrt((double)f[i]));
        grid[i] = (uint8_t)(q > 127.0 ? 127.0 : q);
    }
}

int monte_mc_clearance_grid(const monte_mc_volume *vol, const uint8_t *labels, int n_materials, int heavy_material,
                            int cell_log2, uint8_t *grid) {
    int32_t d[3];
    if (int rc = monte_mc_clearance_dims(vol, cell_log2, d)) return rc;
    if (!labels || !grid || n_materials < 1 || heavy_material < 0 || heavy_material >= n_materials) {
        monte::set_error("monte_mc_clearance_grid: bad argument");
        return MONTE_E_ARG;
    }
    std::vector<float> f((size_t)d[0] * d[1] * d[2]);
    clearance_seed(vol, labels, n_materials, heavy_material, cell_log2, d[0], d[1], f);
    const int sign[3] = {0, 0, 0};
    clearance_transform(f, d[0], d[1], d[2], sign, grid);
    return MONTE_OK;
}

/* Directional form (MONTE_MC_TRACK_DIRECTIONAL): eight grids, one per octant of the flight direction, index
 * o = (dx > 0) | (dy > 0) << 1 | (dz > 0) << 2, laid out [o][cz][cy][cx].  grid[o][cell] bounds the distance to the
 * heavy material in the cells a ray with that direction can still reach (componentwise at or beyond the cell), so a
 * photon flying away from the dense insert has an unbounded clearance.  Octants are transformed concurrently.     */
int monte_mc_clearance_grid_octants(const monte_mc_volume *vol, const uint8_t *labels, int n_materials, int heavy_material,
                                    int cell_log2, uint8_t *grid8) {
    int32_t d[3];
    if (int rc = monte_mc_clearance_dims(vol, cell_log2, d)) return rc;
    if (!labels || !grid8 || n_materials < 1 || heavy_material < 0 || heavy_material >= n_materials) {
        monte::set_error("monte_mc_clearance_grid_octants: bad argument");
        return MONTE_E_ARG;
    }
    const size_t ncell = (size_t)d[0] * d[1] * d[2];
    std::vector<float> seed(ncell);
    clearance_seed(vol, labels, n_materials, heavy_material, cell_log2, d[0], d[1], seed);
    std::vector<std::thread> th;
    for (int o = 0; o < 8; o++)
        th.emplace_back([&, o] {
            std::vector<float> f(seed);
            const int sign[3] = {(o & 1) ? 1 : -1, (o & 2) ? 1 : -1, (o & 4) ? 1 : -1};
            clearance_transform(f, d[0], d[1], d[2], sign, grid8 + (size_t)o * ncell);
        });
    for (auto &t : th) t.join();
    return MONTE_OK;
}

/* the material whose attenuation sets the majorant: argmax of total*density at 60 keV (-1: fewer than 2 materials) */
int monte_xs_heavy_material(const monte_mc_xs *xs) {
    if (!xs || xs->n_materials < 2 || xs->n_materials > MONTE_MC_MAX_MATERIALS) return -1;
    int best = 0;
    for (int m = 1; m < xs->n_materials; m++)
        if (xs->total[m][60] * xs->density[m] > xs->total[best][60] * xs->density[best]) best = m;
    return best;
}

int monte_mc_resolve_tracking(const monte_mc_xs *xs, const monte_mc_spectrum *spec, int32_t *cell_log2, double *ratio) {
    if (cell_log2) *cell_log2 = 2;
    if (ratio) *ratio = 1.0;
    const int heavy = monte_xs_heavy_material(xs);
    if (heavy < 0) return MONTE_MC_TRACK_GLOBAL;
    auto r_at = [&](double keV) {
        int k = (int)(keV + 0.5);
        if (k < 1) k = 1;
        if (k > MONTE_MC_TABLE_ROWS - 1) k = MONTE_MC_TABLE_ROWS - 1;
        double mx = 0, lo = 0;
        for (int m = 0; m < xs->n_materials; m++) {
            const double v = (double)xs->total[m][k] * (double)xs->density[m];
            if (v > mx) mx = v;
            if (m != heavy && v > lo) lo = v;
        }
        return lo > 0 ? mx / lo : 1.0;
    };
    double mean = 1.0;
    if (spec && spec->n_bins > 0 && spec->cdf) {             // weights = the bin probabilities of the CDF
        double acc = 0, wsum = 0;
        for (int b = 0; b < spec->n_bins; b++) {
            const double w = (double)spec->cdf[b + 1] - (double)spec->cdf[b];
            if (w <= 0) continue;
            acc += w * r_at((b + 1) * spec->bin_keV);
            wsum += w;
        }
        if (wsum > 0) mean = acc / wsum;
    } else mean = r_at(spec ? spec->mono_keV : 140.0);
    if (ratio) *ratio = mean;
    return mean > 3.0 ? MONTE_MC_TRACK_DIRECTIONAL : MONTE_MC_TRACK_GLOBAL;
}

}  // extern "C"
