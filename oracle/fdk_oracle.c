/* oracle/fdk_oracle.c — CPU restatement of the reference FDK path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (libmonte_gpu) never links, loads or calls it.
 *
 * Follows, statement by statement and in double precision exactly as written there,
 *   recon/bp3d20.cpp:36-43     step 1  cosine weight            -> oracle_fdk_weight
 *   recon/bp3d20.cpp:48-60     Ram-Lak taps (float)             -> oracle_fdk_ramp
 *   recon/bp3d20.cpp:63-73     step 2  convolution + transpose  -> oracle_fdk_filter
 *   recon/bp3d20.cpp:83-166    step 3  backprojection           -> oracle_fdk_backproject
 *   recon/bp3d20_325.cpp       same with 0.1 cm pitch, no sphere mask, extra *5 (:134-135,170)
 *   recon/fbp2.cpp:36-152      2-D fan-beam variant             -> oracle_fbp2
 * with the literals lifted into monte_fdk_geom (include/monte_gpu.h).
 *
 * Pinning: tests/test_oracle_fdk.py runs the UNMODIFIED reference binaries built into
 * oracle/_ref (oracle/Makefile) on seeded inputs and requires bit equality of the filtered
 * maps and of the reconstructed slab; the same outputs are committed (sub-sampled + sha256)
 * under tests/golden/ for boxes without /root/reference.
 *
 * One deliberate definition: the reference's bilinear fetch reads up to nu+1 floats past the
 * end of map_out for the last view (bp3d20.cpp:152-156; heap over-read).  Here and in the CUDA
 * path those elements are 0.0f (what a fresh calloc'd mmap chunk holds in practice).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../include/monte_gpu.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static double view_angle(const monte_fdk_geom *g, int v) {
    return g->angle0_deg + g->angle_step_deg * (double)v;
}

/* bp3d20.cpp:36-43.  map[v][zeta][p] -> map_w, float result of a double product. */
void oracle_fdk_weight(const monte_fdk_geom *g, const float *map, float *map_w) {
    const int nu = g->nu, nv = g->nv;
    const double wd = (g->weight_mode == MONTE_FDK_TEXTBOOK) ? g->dsd : g->weight_dist;
    for (int v = 0; v < g->n_views; v++)
        for (int zeta = 0; zeta < nu; zeta++)
            for (int p = 0; p < nv; p++) {
                size_t i = (size_t)nu * nv * v + (size_t)nv * zeta + p;
                double a = -1 * (zeta * g->du) + g->half_u;
                double b = (p * g->dv) - g->half_v;
                map_w[i] = map[i] * (wd / sqrt(pow(wd, 2) + pow(a, 2) + pow(b, 2)));
            }
}

/* bp3d20.cpp:48-60.  ramp has 2*nu-1 floats, centre at nu-1. */
void oracle_fdk_ramp(int nu, float *ramp) {
    memset(ramp, 0, sizeof(float) * (size_t)(2 * nu - 1));
    ramp[nu - 1] = 0.25;
    for (int n = 1; n < nu; n++) {
        if (n % 2 == 0) {
            ramp[nu - 1 + n] = 0;
            ramp[nu - 1 - n] = 0;
        } else {
            ramp[nu - 1 + n] = -1. / pow(n * M_PI, 2);
            ramp[nu - 1 - n] = -1. / pow(n * M_PI, 2);
        }
    }
}

/* bp3d20.cpp:63-73.  out[v][d][b] = sum_c map_w[v][c][d]*scale*ramp[nu-1-b+c], float accumulator,
 * each term formed in double. */
void oracle_fdk_filter(const monte_fdk_geom *g, const float *map_w, float *out) {
    const int nu = g->nu, nv = g->nv;
    float *ramp = (float *)malloc(sizeof(float) * (size_t)(2 * nu - 1));
    oracle_fdk_ramp(nu, ramp);
    /* TEXTBOOK: taps of the continuous Ram-Lak at the iso-centre pitch du*Dso/Dsd */
    const double scale = (g->weight_mode == MONTE_FDK_TEXTBOOK) ? g->dsd / (g->dso * g->du) : g->filter_scale;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < g->n_views; v++)
        for (int d = 0; d < nv; d++)
            for (int b = 0; b < nu; b++) {
                float tmp = 0.;
                for (int c = 0; c < nu; c++)
                    tmp += map_w[(size_t)v * nu * nv + (size_t)c * nv + d] * scale * ramp[nu - 1 - b + c];
                out[(size_t)v * nu * nv + (size_t)d * nu + b] = tmp;
            }
    free(ramp);
}

static inline float fetch(const float *out, size_t total, size_t i) {
    return i < total ? out[i] : 0.0f;
}

/* bp3d20.cpp:83-166.  vol_xy[z][t][s] (and vol_zy[s][t][z] if non-NULL) are ACCUMULATED into,
 * views outermost in the reference; here voxels are the parallel axis and views the inner loop,
 * which gives every voxel the same sequence of float additions.                                 */
void oracle_fdk_backproject(const monte_fdk_geom *g, const float *filt, float *vol_xy, float *vol_zy) {
    const int nu = g->nu, nv = g->nv, nviews = g->n_views;
    const size_t total = (size_t)nviews * nu * nv;
    const int textbook = (g->weight_mode == MONTE_FDK_TEXTBOOK);
    const double inv_du = 1.0 / g->du, inv_dv = 1.0 / g->dv;
    /* the reference writes the scale as an integer literal where it is one (-2*, *10.) */
    double *cb = (double *)malloc(sizeof(double) * nviews * 6);
    for (int v = 0; v < nviews; v++) {
        double beta = view_angle(g, v);
        float start_x = (float)(-g->dso), start_y = 0;
        double c = cos(M_PI * beta / 180), s = sin(M_PI * beta / 180);
        cb[6 * v + 0] = c;
        cb[6 * v + 1] = s;
        cb[6 * v + 2] = start_x * c - start_y * s;              /* primary_x  :87 */
        cb[6 * v + 3] = start_x * s + start_y * c;              /* primary_y  :88 */
        cb[6 * v + 4] = cos(-1 * M_PI * beta / 180);
        cb[6 * v + 5] = sin(-1 * M_PI * beta / 180);
    }
    const double beta_span = (float)g->angle_step_deg;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int z = g->z_begin; z < g->z_end; z++)
        for (int t = g->t_begin; t < g->t_end; t++)
            for (int s = g->s_begin; s < g->s_end; s++) {
                if (g->mask_r2 >= 0) {
                    long long dz = z - g->mask_cz, dt = t - g->mask_ct, ds = s - g->mask_cs;
                    if (dz * dz + dt * dt + ds * ds > g->mask_r2) continue;
                }
                float acc = vol_xy[(size_t)g->ny * g->nx * z + (size_t)g->nx * t + s];
                float acc_zy = vol_zy ? vol_zy[(size_t)g->ny * g->nz * s + (size_t)g->nz * t + z] : 0.f;
                for (int v = 0; v < nviews; v++) {
                    const double beta = view_angle(g, v);
                    const double cB = cb[6 * v], sB = cb[6 * v + 1];
                    double pv0 = (g->x0 + s * g->vox) - cb[6 * v + 2];
                    double pv1 = (g->y0 - t * g->vox) - cb[6 * v + 3];
                    double pv2 = (g->z0 - z * g->vox);
                    double tmp_x = pv0;
                    pv0 = pv0 * cb[6 * v + 4] - pv1 * cb[6 * v + 5];
                    pv1 = tmp_x * cb[6 * v + 5] + pv1 * cb[6 * v + 4];
                    double to_det = g->dsd / pv0;
                    pv0 *= to_det; pv1 *= to_det; pv2 *= to_det;
                    if (fabs(pv1) > g->half_u || fabs(pv2) > g->half_v) continue;
                    double y, x;
                    if (g->coord_mode == MONTE_FDK_COORD_SCALE_BEFORE) {
                        y = -1 * (pv1 * inv_du - nu / 2.);
                        x = -1 * (pv2 * inv_dv - nv / 2.);
                    } else {
                        y = -inv_du * (pv1 - g->half_u);
                        x = -inv_dv * (pv2 - g->half_v);
                    }
                    int yi = (int)y, xi = (int)x;
                    double wgt;
                    if (!textbook) {
                        double tmp_s = (g->x0 + s * g->vox) * cB - (g->y0 - t * g->vox) * sB;
                        double tmp_t = (g->x0 + s * g->vox) * sB + (g->y0 - t * g->vox) * cB;
                        double d = fabs(-1 * tan(beta) * tmp_t + 1 * tmp_s) / (sqrt(1 + pow(tan(beta), 2)));
                        if (tmp_s < 0) d *= -1;
                        wgt = pow(g->weight_dist, 2) / pow(g->weight_dist - d, 2);
                    } else {
                        /* U = distance of the voxel from the source along the central ray */
                        double U = ((g->x0 + s * g->vox) * cB + (g->y0 - t * g->vox) * sB) + g->dso;
                        wgt = pow(g->dso, 2) / pow(U, 2);
                    }
                    double tmp_output = 0;
                    if (0 <= x && x <= nv && 0 <= y && y <= nu) {
                        size_t base = (size_t)v * nu * nv;
                        double geo_bl =
                            (yi + 1 - y) * ((xi + 1 - x) * fetch(filt, total, base + (size_t)xi * nu + yi) +
                                            (x - xi) * fetch(filt, total, base + (size_t)(xi + 1) * nu + yi)) +
                            (y - yi) * ((xi + 1 - x) * fetch(filt, total, base + (size_t)xi * nu + yi + 1) +
                                        (x - xi) * fetch(filt, total, base + (size_t)(xi + 1) * nu + yi + 1));
                        tmp_output = geo_bl;
                    }
                    double output = wgt * tmp_output * beta_span * 2 * M_PI / 360;
                    if (textbook) {                 /* f = 1/2 * sum (Dso/U)^2 * p~ * dbeta */
                        acc += output * 0.5;
                        acc_zy += output * 0.5;
                    } else {
                        acc += output * g->out_scale * g->out_scale2;
                        acc_zy += output * g->out_scale * g->out_scale2;
                    }
                }
                vol_xy[(size_t)g->ny * g->nx * z + (size_t)g->nx * t + s] = acc;
                if (vol_zy) vol_zy[(size_t)g->ny * g->nz * s + (size_t)g->nz * t + z] = acc_zy;
            }
    free(cb);
}

/* whole pipeline, buffers as in monte_gpu_fdk (volumes are zeroed first) */
int oracle_fdk(const monte_fdk_geom *g, const float *map, float *filtered, float *vol_xy, float *vol_zy) {
    size_t n = (size_t)g->n_views * g->nu * g->nv;
    float *mw = (float *)malloc(sizeof(float) * n);
    float *f = filtered ? filtered : (float *)malloc(sizeof(float) * n);
    if (!mw || !f) return -1;
    oracle_fdk_weight(g, map, mw);
    oracle_fdk_filter(g, mw, f);
    memset(vol_xy, 0, sizeof(float) * (size_t)g->nx * g->ny * g->nz);
    if (vol_zy) memset(vol_zy, 0, sizeof(float) * (size_t)g->nx * g->ny * g->nz);
    oracle_fdk_backproject(g, f, vol_xy, vol_zy);
    free(mw);
    if (!filtered) free(f);
    return 0;
}

/* recon/fbp2.cpp:36-152.  sino[v][nu]; image[t][s] (image_xy[outH*t+s]). */
int oracle_fbp2(const monte_fdk_geom *g, int view_first, const float *sino, float *filtered, float *image) {
    const int nu = g->nu, nviews = g->n_views;
    float *pw = (float *)malloc(sizeof(float) * (size_t)nu * nviews);
    double *ramp = (double *)calloc((size_t)(2 * nu - 1), sizeof(double));
    const double wd = g->weight_dist;
    for (int b = 0; b < nviews; b++)
        for (int zeta = 0; zeta < nu; zeta++)   /* fbp2.cpp:38 */
            pw[b * nu + zeta] = sino[b * nu + zeta] * (wd / (sqrt(pow(wd, 2) + pow(-g->half_u + g->du * zeta, 2))));
    ramp[nu - 1] = 0.25;
    for (int n = 1; n < nu; n++) {
        if (n % 2 == 0) { ramp[nu - 1 + n] = 0; ramp[nu - 1 - n] = 0; }
        else { ramp[nu - 1 + n] = -1. / pow(n * M_PI, 2); ramp[nu - 1 - n] = -1. / pow(n * M_PI, 2); }
    }
    for (int d = 0; d < nviews; d++)            /* fbp2.cpp:63-76 (double accumulator) */
        for (int b = 0; b < nu; b++) {
            double tmp = 0.;
            for (int c = 0; c < nu; c++) tmp += pw[d * nu + c] * g->filter_scale * ramp[nu - 1 - b + c];
            filtered[d * nu + b] = tmp;
        }
    memset(image, 0, sizeof(float) * (size_t)g->nx * g->ny);
    const double inv_du = 1.0 / g->du;
    const double beta_span = (float)g->angle_step_deg;
    for (int v = view_first; v < nviews; v++) {   /* fbp2.cpp:89 */
        double beta = view_angle(g, v);
        float start_x = (float)(-g->dso), start_y = 0;
        double primary_x = start_x * cos(M_PI * beta / 180) - start_y * sin(M_PI * beta / 180);
        double primary_y = start_x * sin(M_PI * beta / 180) + start_y * cos(M_PI * beta / 180);
        for (int t = g->t_begin; t < g->t_end; t++)
            for (int s = g->s_begin; s < g->s_end; s++) {
                double pv0 = (g->x0 + s * g->vox) - primary_x;
                double pv1 = (g->y0 - t * g->vox) - primary_y;
                double tmp_x = pv0;
                pv0 = pv0 * cos(-1 * M_PI * beta / 180) - pv1 * sin(-1 * M_PI * beta / 180);
                pv1 = tmp_x * sin(-1 * M_PI * beta / 180) + pv1 * cos(-1 * M_PI * beta / 180);
                double to_det = g->dsd / pv0;
                pv0 *= to_det; pv1 *= to_det;
                if (fabs(pv1) > g->half_u) continue;
                int index_y = -inv_du * ((pv1 - g->half_u));      /* fbp2.cpp:126 */
                double tmp_s = (g->x0 + s * g->vox) * cos(M_PI * beta / 180) - (g->y0 - g->vox * t) * sin(M_PI * beta / 180);
                double tmp_t = (g->x0 + s * g->vox) * sin(M_PI * beta / 180) + (g->y0 - g->vox * t) * cos(M_PI * beta / 180);
                double d = fabs(-1 * tan(beta) * tmp_t + 1 * tmp_s) / (sqrt(1 + pow(tan(beta), 2)));
                if (tmp_s < 0) d *= -1;
                size_t fi = (size_t)v * nu + index_y;
                float fv = fi < (size_t)nu * nviews ? filtered[fi] : 0.f;
                double tmp_output = (pow(wd, 2) / pow(wd - d, 2)) * fv * beta_span * 2 * M_PI / 360;
                image[(size_t)g->nx * t + s] += tmp_output * g->out_scale;
            }
    }
    free(pw); free(ramp);
    return 0;
}

/* presets (same literals as libmonte_gpu's monte_fdk_geom_* — duplicated on purpose so the
 * oracle does not depend on the product library) */
static void preset_common(monte_fdk_geom *g) {
    memset(g, 0, sizeof(*g));
    g->n_views = 360;
    g->half_u = g->half_v = 16.25;
    g->dso = 160; g->dsd = 220; g->weight_dist = 60;
    g->filter_scale = 0.5; g->out_scale = 2.7; g->out_scale2 = 1;
    g->angle0_deg = 0; g->angle_step_deg = 1;
    g->nx = g->ny = g->nz = 256; g->vox = 0.1;
    g->x0 = -12.8; g->y0 = 12.8; g->z0 = 12.8;
    g->s_begin = 125; g->s_end = 130; g->t_begin = 0; g->t_end = 256; g->z_begin = 0; g->z_end = 256;
    g->mask_r2 = -1;
    g->weight_mode = MONTE_FDK_REFERENCE;
}
void oracle_fdk_geom_bp3d20(monte_fdk_geom *g) {
    preset_common(g);
    g->nu = g->nv = 65; g->du = g->dv = 0.5;
    g->mask_cs = g->mask_ct = g->mask_cz = 128; g->mask_r2 = 118 * 118;
    g->coord_mode = MONTE_FDK_COORD_SCALE_AFTER;
}
void oracle_fdk_geom_bp3d20_325(monte_fdk_geom *g) {
    preset_common(g);
    g->nu = g->nv = 325; g->du = g->dv = 0.1;
    g->out_scale2 = 5;
    g->coord_mode = MONTE_FDK_COORD_SCALE_BEFORE;
}
void oracle_fdk_geom_fbp2(monte_fdk_geom *g) {
    preset_common(g);
    g->nu = 65; g->nv = 1; g->du = 0.5; g->dv = 0.5;
    g->nz = 1; g->z_end = 1; g->s_begin = 0; g->s_end = 256;
    g->out_scale = 1.7;
}
