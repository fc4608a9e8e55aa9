/* oracle/mc_oracle.c — CPU restatement of the reference photon-history Monte Carlo.
 * TEST INFRASTRUCTURE ONLY (see fdk_oracle.c header for who may load it).
 *
 * Follows monte_cpp/CBCT_real2.cpp (double precision, MT19937, the CPU form) and
 * monte_cu/CBCT_real325im.cu (voxel labels, several materials, com_flag guard, tallies image0 /
 * image5) — each block cites the lines it restates.  The flow of one history is the reference's:
 *   source set-up            CBCT_real2.cpp:205-266      CBCT_real325im.cu:464-540
 *   delta_sampling           CBCT_real2.cpp:733-815      CBCT_real325im.cu:862-976
 *   primary detection        CBCT_real2.cpp:283-327      CBCT_real325im.cu:552-590
 *   interaction loop         CBCT_real2.cpp:337-588      CBCT_real325im.cu:599-845
 *   Kahn Compton sampler     CBCT_real2.cpp:404-448      CBCT_real325im.cu:701-757
 *   direction update         CBCT_real2.cpp:460-470      CBCT_real325im.cu:768-780
 *   scatter detection        CBCT_real2.cpp:528-564      CBCT_real325im.cu:823-843
 *
 * Two configurations of the SAME code:
 *  (1) quirks = ORACLE_Q_REAL2: every behaviour-changing quirk of the shipped CBCT_real2.cpp
 *      switched on.  With the reference's MT19937 seeded identically (oracle/shim/windows.h pins
 *      its time() seed) this reproduces the unmodified binary's output image and counters BIT FOR
 *      BIT (tests/test_oracle_mc.py, golden in tests/golden/mc_real2.npz).  This is what pins
 *      the restatement of every shared routine (Woodcock loop, Kahn sampler, angle update,
 *      detector binning).
 *  (2) quirks = 0: the documented intended physics ("Recommended"): exact aim at
 *      the pixel, analytic flight through air, per-voxel material, majorant over all materials,
 *      float energies, com_flag guard, symmetric detector bounds.  The CUDA kernel implements
 *      exactly this; with rng_mode = ORACLE_RNG_PHILOX both draw the same Philox2x32-10 variates
 *      per history, so they can be compared history by history, not only statistically.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/monte_gpu.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------------------------------------------------------------- quirk bits */
#define OQ_EXTRA_DRAW        (1 << 0)  /* one unused variate per history     CBCT_real2.cpp:176     */
#define OQ_PIXEL_OFFSET      (1 << 1)  /* yl = (half-pixel/2) - pixel*i       CBCT_real2.cpp:210     */
#define OQ_POLAR_ATAN        (1 << 2)  /* theta_a = pi/2 - atan(zl/Dsd)       CBCT_real2.cpp:213     */
#define OQ_PREADVANCE        (1 << 3)  /* start at fraction (Dso-start)/Dsd, Woodcock through air
                                          until |x|>=62,|y|>=62,|z|>=17       CBCT_real2.cpp:254,782 */
#define OQ_MUMAX_FIRST       (1 << 4)  /* majorant = mu of material 1 only    CBCT_real2.cpp:736     */
#define OQ_NU_FLOAT          (1 << 5)  /* acceptance variate stored as float  CBCT_real2.cpp:761     */
#define OQ_EXIT_TEST_BUG     (1 << 6)  /* y test uses the x formula, phi=0    CBCT_real2.cpp:343-345 */
#define OQ_FIRST_MATERIAL    (1 << 7)  /* label==1 is material 1, anything else air; tables of
                                          material 1 at every site            CBCT_real2.cpp:352-356,771 */
#define OQ_ENERGY_INT        (1 << 8)  /* `int Energy`                        CBCT_real2.cpp:181,446 */
#define OQ_FIRST_COMPTON_Z   (1 << 9)  /* _new copied unconditionally         CBCT_real2.cpp:460-463 */
#define OQ_PHI_NEG           (1 << 10) /* phi = -2 pi u                       CBCT_real2.cpp:451     */
#define OQ_NO_PRIMARY_TALLY  (1 << 11) /* primaries tallied only when q==0    CBCT_real2.cpp:290     */
#define OQ_DETECT_UNROT      (1 << 12) /* d from un-rotated coordinates, view index as radians,
                                          strict |d|<half                     CBCT_real2.cpp:530-544 */
#define OQ_LABEL_RINT        (1 << 13) /* voxel index rint(p*10)+c, strict clip box  CBCT_real2.cpp:770-771 */
#define OQ_PRIMARY_STRICT    (1 << 14) /* primary iff rotated x > Dod         CBCT_real2.cpp:287     */
#define ORACLE_Q_REAL2       ((1 << 15) - 1)

#define ORACLE_RNG_MT     0
#define ORACLE_RNG_PHILOX 1

typedef struct oracle_mc_tables {       /* double tables, index = keV 0..200, [material][row] */
    int32_t n_materials;
    double density[MONTE_MC_MAX_MATERIALS];
    double coh[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];
    double compt[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];
    double photo[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];
    double total[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];
    /* Rayleigh form-factor tables (monte_mc_xs.ff_*; used only with coherent_mode FORMFACTOR, which the
     * reference does not have: its coherent event keeps the direction, CBCT_real325im.cu:656-695) */
    int32_t ff_points;
    double ff_x2[MONTE_MC_MAX_MATERIALS][MONTE_MC_FF_POINTS];
    double ff_cum[MONTE_MC_MAX_MATERIALS][MONTE_MC_FF_POINTS];
} oracle_mc_tables;

typedef struct oracle_mc_opts {
    int32_t  rng_mode;
    int32_t  quirks;
    uint64_t seed;
    double   preadvance_start;   /* 6 (CBCT_real2.cpp:254) or 10.1 (CBCT_real325im.cu:460) */
    double   track_box[3];       /* 62, 62, 17 */
    int32_t  n_threads;          /* 0 = all (MT mode with >1 thread is not the reference's stream) */
    double   label_center[3];    /* OQ_LABEL_RINT: 90, 90, 160 (CBCT_real2.cpp:771) */
    /* monte_mc_volume.tracking_mode == MONTE_MC_TRACK_CLEARANCE (not in the reference): the grid of
     * monte_mc_clearance_grid (include/monte_gpu.h) and the material it was built for */
    const uint8_t *clear_grid;
    int32_t  clear_dims[3];
    int32_t  heavy;
} oracle_mc_opts;

typedef struct oracle_mc_result {
    uint64_t histories, primaries, scatter_detected, absorbed, interactions, coherent, compton;
    uint64_t woodcock_steps;
    int64_t  num_scatter;        /* CBCT_real2.cpp:332,377: non-primary minus coherent events */
    uint64_t num_nd;             /* CBCT_real2.cpp:580-583 */
    double   sum_e_primary, sum_e_scatter;
} oracle_mc_result;

/* ---------------------------------------------------------------- MT19937 (mt19937ar algorithm,
 * Matsumoto & Nishimura 2002; the generator bundled as monte_cpp/Mersenne_twister.cpp) */
typedef struct { uint32_t mt[624]; int mti; } mt_state;
static void mt_seed(mt_state *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->mti = 624;
}
static uint32_t mt_u32(mt_state *s) {
    if (s->mti >= 624) {
        uint32_t *m = s->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (m[k] & 0x80000000u) | (m[(k + 1) % 624] & 0x7fffffffu);
            m[k] = m[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->mti = 0;
    }
    uint32_t y = s->mt[s->mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* genrand_real3, Mersenne_twister.cpp:117-121 */
static double mt_real3(mt_state *s) { return (((double)mt_u32(s)) + 0.5) * (1.0 / 4294967296.0); }

/* ---------------------------------------------------------------- Philox2x32-10 (Salmon et al., SC'11) */
static void philox2x32(uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += 0x9E3779B9u;
    }
    out[0] = c0; out[1] = c1;
}
void oracle_philox2x32(uint32_t c0, uint32_t c1, uint32_t key, uint32_t *out) { philox2x32(c0, c1, key, out); }

/* 23-bit uniform in (0,1), exactly representable in fp32 (the CUDA kernel uses the same map) */
static double u01_23(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }

#define STREAM_FLIGHT 0u
#define STREAM_EVENT  1u
#define STREAM_SOURCE 2u

typedef struct {
    int mode;
    mt_state *mt;
    uint32_t key, c0, c1hi;       /* Philox: per-history counter words */
    uint32_t n_flight, n_event;
} rng_t;

static void rng_history(rng_t *r, uint64_t seed, uint64_t hid) {
    r->key = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u);
    r->c0 = (uint32_t)hid;
    r->c1hi = ((uint32_t)(hid >> 32) & 0xFFu) << 24;
    r->n_flight = r->n_event = 0;
}
static void rng_pair(rng_t *r, uint32_t stream, uint32_t idx, double *a, double *b) {
    uint32_t o[2];
    philox2x32(r->c0, r->c1hi | (stream << 22) | (idx & 0x3FFFFFu), r->key, o);
    *a = u01_23(o[0]); *b = u01_23(o[1]);
}
/* one Woodcock step needs (beta, nu) in this order (CBCT_real2.cpp:750,761) */
static void draw_step(rng_t *r, double *beta, double *nu) {
    if (r->mode == ORACLE_RNG_MT) { *beta = mt_real3(r->mt); *nu = mt_real3(r->mt); }
    else rng_pair(r, STREAM_FLIGHT, r->n_flight++, beta, nu);
}

/* ---------------------------------------------------------------- scene */
typedef struct {
    const monte_mc_geom *g;
    const monte_mc_volume *vol;
    const uint8_t *labels;
    const oracle_mc_tables *tb;
    const monte_mc_spectrum *spec;
    const oracle_mc_opts *o;
    uint32_t present;            /* materials the majorant is taken over (bit m); all ones unless
                                    monte_mc_volume.majorant_mode == MONTE_MC_MAJORANT_PRESENT (not in the reference,
                                    which takes every table it loaded, CBCT_real325im.cu:866-868) */
} scene_t;

typedef struct {                 /* class photon, CBCT_real2.cpp:38-49 (fields that matter) */
    double x, y, z, x_p, y_p, z_p, length;
    int at_entry;                /* tracking_mode CLEARANCE: the photon sits on the face of the clip box it entered through */
} photon_t;

static int table_index(double E) {
    int k = (int)(E + 0.5);
    if (k < 0) k = 0;
    if (k > MONTE_MC_TABLE_ROWS - 1) k = MONTE_MC_TABLE_ROWS - 1;
    return k;
}
static int label_material(const scene_t *S, int label) {   /* label -> table row, -1 = air */
    if (S->o->quirks & OQ_FIRST_MATERIAL) return label == 1 ? 0 : -1;
    if (label == 0) return -1;
    return label <= S->tb->n_materials ? label - 1 : S->tb->n_materials - 1;
}
static double mu_max_at(const scene_t *S, int k) {
    if (S->o->quirks & OQ_MUMAX_FIRST) return S->tb->total[0][k] * S->tb->density[0];
    double m = 0;
    for (int i = 0; i < S->tb->n_materials; i++) {
        if (!(S->present >> i & 1u)) continue;
        double v = S->tb->total[i][k] * S->tb->density[i];
        if (v > m) m = v;
    }
    return m;
}

/* majorant over every material but the heavy one (tracking_mode CLEARANCE) */
static double mu_lo_at(const scene_t *S, int k) {
    double m = 0;
    for (int i = 0; i < S->tb->n_materials; i++) {
        if (i == S->o->heavy || !(S->present >> i & 1u)) continue;
        double v = S->tb->total[i][k] * S->tb->density[i];
        if (v > m) m = v;
    }
    return m;
}
/* clearance (cm) of the cell holding (x,y,z): no voxel of the heavy material within this distance -- for
 * tracking_mode DIRECTIONAL, none that a ray with direction (dx,dy,dz) can still reach (eight grids, one per octant) */
static double clearance_at(const scene_t *S, double x, double y, double z, double dx, double dy, double dz) {
    const monte_mc_volume *v = S->vol;
    const double inv = 1.0 / v->pitch;
    int ix = (int)floor((x - v->origin[0]) * inv), iy = (int)floor((y - v->origin[1]) * inv), iz = (int)floor((z - v->origin[2]) * inv);
    if (ix < 0) ix = 0; if (iy < 0) iy = 0; if (iz < 0) iz = 0;
    if (ix > v->nx - 1) ix = v->nx - 1; if (iy > v->ny - 1) iy = v->ny - 1; if (iz > v->nz - 1) iz = v->nz - 1;
    const int s = v->clearance_cell_log2;
    const size_t ncell = (size_t)S->o->clear_dims[0] * S->o->clear_dims[1] * S->o->clear_dims[2];
    const size_t oct = v->tracking_mode == MONTE_MC_TRACK_DIRECTIONAL ? (size_t)((dx > 0) | ((dy > 0) << 1) | ((dz > 0) << 2)) : 0;
    const int q = S->o->clear_grid[oct * ncell + ((size_t)(iz >> s) * S->o->clear_dims[1] + (iy >> s)) * S->o->clear_dims[0] + (ix >> s)];
    return q * (0.5 * (double)(1 << s) * v->pitch);
}

/* voxel label at (x,y,z), 0 outside the clip box */
static int lookup(const scene_t *S, double x, double y, double z) {
    const monte_mc_volume *v = S->vol;
    if (S->o->quirks & OQ_LABEL_RINT) {
        /* CBCT_real2.cpp:770-771: strict box, geometry[int(nx*ny*rint(z*10+cz) + nx*(rint(y*10)+cy) + (rint(x*10)+cx))] */
        if (!(v->clip_lo[0] < x && x < v->clip_hi[0] && v->clip_lo[1] < y && y < v->clip_hi[1] &&
              v->clip_lo[2] < z && z < v->clip_hi[2])) return 0;
        const double inv = 1.0 / v->pitch;
        const double cx = S->o->label_center[0], cy = S->o->label_center[1], cz = S->o->label_center[2];
        long long idx = (long long)((double)v->nx * v->ny * (rint(z * inv + cz)) + (double)v->nx * (rint(y * inv) + cy) + (rint(x * inv) + cx));
        if (idx < 0 || idx >= (long long)v->nx * v->ny * v->nz) return 0;
        return S->labels[idx];
    }
    if (!(v->clip_lo[0] <= x && x < v->clip_hi[0] && v->clip_lo[1] <= y && y < v->clip_hi[1] &&
          v->clip_lo[2] <= z && z < v->clip_hi[2])) return 0;
    const double inv = 1.0 / v->pitch;
    int ix = (int)floor((x - v->origin[0]) * inv), iy = (int)floor((y - v->origin[1]) * inv), iz = (int)floor((z - v->origin[2]) * inv);
    if (ix < 0 || iy < 0 || iz < 0 || ix >= v->nx || iy >= v->ny || iz >= v->nz) return 0;
    return S->labels[((size_t)iz * v->ny + iy) * v->nx + ix];
}

/* delta_sampling, CBCT_real2.cpp:733-815 / CBCT_real325im.cu:862-976.
 * Returns 1 if the photon stopped at a real collision, 0 if it left the tracking region.
 * In quirk mode the tracking region is the reference's air box; otherwise it is the clip box
 * (air outside it is crossed analytically, which is the same distribution: every tentative
 * collision in air is rejected).                                                               */
static int delta_sampling(const scene_t *S, rng_t *R, photon_t *p, double E, double sin_theta_a,
                          double cos_theta_a, double sin_phi_a, double cos_phi_a, uint64_t *steps) {
    const int k = table_index(E);
    const double mu_max = mu_max_at(S, k);
    double x = p->x, y = p->y, z = p->z, length = 0;
    p->x_p = p->x; p->y_p = p->y; p->z_p = p->z;
    const int q = S->o->quirks;
    const monte_mc_volume *v = S->vol;
    int collided = 0;
    const int clearance = (v->tracking_mode == MONTE_MC_TRACK_CLEARANCE || v->tracking_mode == MONTE_MC_TRACK_ADAPTIVE ||
                           v->tracking_mode == MONTE_MC_TRACK_DIRECTIONAL) && S->o->clear_grid && !q;
    const double mu_lo = clearance ? mu_lo_at(S, k) : 0;
    for (;;) {
        double beta, nu;
        draw_step(R, &beta, &nu);
        if (clearance && mu_lo > 0) {
            /* two-level majorant (not in the reference; same distribution of collision sites): inside the
             * clearance radius D of the current cell only the lighter materials occur, so the flight is sampled
             * with their majorant; a flight longer than D stops at D without a collision (memoryless) */
            /* on the entry face the cell is taken 1e-3 of a voxel side further along the ray (as in the kernel) */
            const double nud = p->at_entry ? 1e-3 * v->pitch : 0.0;
            p->at_entry = 0;
            const double ux = sin_theta_a * cos_phi_a, uy = sin_theta_a * sin_phi_a, uz = cos_theta_a;
            const double D = clearance_at(S, x + nud * ux, y + nud * uy, z + nud * uz, (float)ux, (float)uy, (float)uz);
            /* ADAPTIVE: only where stopping at D is less likely than a virtual collision would be */
            double thr = 0;
            if (v->tracking_mode == MONTE_MC_TRACK_ADAPTIVE || v->tracking_mode == MONTE_MC_TRACK_DIRECTIONAL) thr = mu_lo < mu_max ? (double)(float)(-log(1.0 - mu_lo / mu_max) / mu_lo) : 1e30;
            if (D > thr) {
                double r = -log(beta) / mu_lo;
                const int cut = r > D;
                if (cut) r = D;
                x += r * sin_theta_a * cos_phi_a; y += r * sin_theta_a * sin_phi_a; z += r * cos_theta_a;
                length += r;
                (*steps)++;
                const int in_box = v->clip_lo[0] <= x && x < v->clip_hi[0] && v->clip_lo[1] <= y && y < v->clip_hi[1] &&
                                   v->clip_lo[2] <= z && z < v->clip_hi[2];
                if (!in_box) {
                    const double far = 1000.0;
                    x += far * sin_theta_a * cos_phi_a; y += far * sin_theta_a * sin_phi_a; z += far * cos_theta_a;
                    break;
                }
                if (cut) continue;
                int m = label_material(S, lookup(S, x, y, z));
                if (m >= 0 && nu <= (S->tb->total[m][k] * S->tb->density[m]) / mu_lo) { collided = 1; break; }
                continue;
            }
        }
        /* mu_max == 0: nothing present attenuates at this energy (an all-air volume under MAJORANT_PRESENT) --
           the medium is transparent, one step of 1e30 cm leaves the tracking region (as in the kernel) */
        double r = mu_max > 0 ? -log(beta) / mu_max : 1e30;
        x += r * sin_theta_a * cos_phi_a;
        y += r * sin_theta_a * sin_phi_a;
        z += r * cos_theta_a;
        length += r;
        (*steps)++;
        if (q & OQ_NU_FLOAT) nu = (float)nu;
        int m = label_material(S, lookup(S, x, y, z));
        if (m < 0) {                         /* air: keep flying, CBCT_real2.cpp:780-785 */
            if (q & OQ_PREADVANCE) {
                if (fabs(x) >= S->o->track_box[0] || fabs(y) >= S->o->track_box[1] || fabs(z) >= S->o->track_box[2]) break;
            } else if (!(v->clip_lo[0] <= x && x < v->clip_hi[0] && v->clip_lo[1] <= y && y < v->clip_hi[1] &&
                         v->clip_lo[2] <= z && z < v->clip_hi[2])) {
                /* left the volume: nothing but air ahead.  Put the end point far along the ray so
                   the reference's two-point detector formulas apply unchanged. */
                const double far = 1000.0;
                x += far * sin_theta_a * cos_phi_a;
                y += far * sin_theta_a * sin_phi_a;
                z += far * cos_theta_a;
                break;
            }
        } else if (nu <= (S->tb->total[m][k] * S->tb->density[m]) / mu_max) {   /* :787-791 */
            collided = 1;
            break;
        }
    }
    p->x = x; p->y = y; p->z = z; p->length = length;
    return collided;
}

/* photon counting (the reference) or energy-integrating response in units of 1/16 keV (monte_gpu.h) */
#define TALLY(x) __atomic_fetch_add(&(x), (g->detector_mode == MONTE_MC_DETECTOR_ENERGY ? (int)(E * MONTE_MC_EID_SCALE + 0.5) : 1), __ATOMIC_RELAXED)

/* detector bin: result = -1*(int(d*inv_pixel - n/2.))   CBCT_real325im.cu:574-575 / CBCT_real2.cpp:311 */
/* ---- Rayleigh angle from a tabulated form factor (not in the reference): x = sin(theta/2)/lambda,
 * x^2 drawn from F(x)^2 on [0, x^2_max = (E/12.398)^2] by inverting the cumulative table, accepted with
 * probability (1 + cos^2 theta)/2, cos theta = 1 - 2 x^2/x^2_max.  Same arithmetic as the CUDA kernel's
 * ray_interp (monte_b200/csrc/mc.cu), in double. */
static double ff_interp(const double *xs, const double *ys, int n, double v) {
    int lo = 0, hi = n - 2;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (xs[mid] <= v) lo = mid; else hi = mid - 1; }
    double w = xs[lo + 1] - xs[lo], t = 0;
    if (w > 0) { t = (v - xs[lo]) / w; if (t < 0) t = 0; if (t > 1) t = 1; }
    return ys[lo] + t * (ys[lo + 1] - ys[lo]);
}
/* one rejection round; returns 1 if accepted */
static int rayleigh_round(const oracle_mc_tables *tb, int m, double E, double r_a, double r_acc, double *cos_theta) {
    const int n = tb->ff_points;
    const double xm = E * (double)(1.0f / 12.3984193f), x2max = xm * xm;
    const double amax = ff_interp(tb->ff_x2[m], tb->ff_cum[m], n, x2max);
    double x2 = ff_interp(tb->ff_cum[m], tb->ff_x2[m], n, r_a * amax);
    if (x2 > x2max) x2 = x2max;
    double c = 1.0 - 2.0 * x2 / x2max;
    if (c < -1) c = -1;
    *cos_theta = c;
    return r_acc <= 0.5 * (1.0 + c * c);
}
int oracle_rayleigh_round(const oracle_mc_tables *tb, int m, double E, double r_a, double r_acc, double *cos_theta) {
    return rayleigh_round(tb, m, E, r_a, r_acc, cos_theta);
}

static int det_bin(double d, double inv_pixel, int n) { return -1 * ((int)(d * inv_pixel - n / 2.)); }

static double sample_energy(const scene_t *S, double u) {
    const monte_mc_spectrum *sp = S->spec;
    if (!sp || sp->n_bins <= 0) return sp ? sp->mono_keV : 140.0;
    double E = sp->mono_keV;                              /* default if not found, CBCT_real325im.cu:491 */
    for (int k = 0; k < sp->n_bins; k++)
        if (sp->cdf[k] <= u && u <= sp->cdf[k + 1]) { E = (k + 1) * sp->bin_keV; break; }   /* :493-497 */
    return E;
}

/* one history.  image0/image5 are [ny][nx] of the current view.  fate (nullable) receives the
 * record described in include/monte_gpu.h (monte_gpu_simulate_fates).                           */
/* Ring detector (monte_gpu.h MONTE_MC_DETECTOR_RING; the geometry of monte_cpp/circle3_2.cpp:162,243-252 generalised to a
 * cylinder about the z axis): where the flight from (xp,yp,zp) through the far point (x,y,z) -- both rotated by -beta --
 * leaves the cylinder of radius ring_radius.  1 and the bin (angle, height) if it does within |z| <= half. */
static int det_bin(double d, double inv_pixel, int n);
static int ring_hit(const monte_mc_geom *g, double xp, double yp, double zp, double x, double y, double z, int *ry, int *rx) {
    const double dx = x - xp, dy = y - yp, dz = z - zp;
    const double R2 = g->ring_radius * g->ring_radius;
    const double a = dx * dx + dy * dy, b = xp * dx + yp * dy, c = xp * xp + yp * yp - R2;
    if (!(a > 0) || x * x + y * y < R2) return 0;
    const double disc = b * b - a * c;
    if (disc < 0) return 0;
    const double t = (sqrt(disc) - b) / a;
    if (!(t > 0)) return 0;
    const double zd = zp + t * dz;
    if (fabs(zd) > g->half) return 0;
    double phi = atan2(yp + t * dy, xp + t * dx);
    if (phi < 0) phi += 2 * M_PI;
    int by = (int)(phi * g->ny / (2 * M_PI));
    if (by > g->ny - 1) by = g->ny - 1;
    *ry = by; *rx = det_bin(zd, 1.0 / g->pixel, g->nx);
    return 1;
}

static void history(const scene_t *S, rng_t *R, int view, int i, int j, uint64_t hid, int32_t *image0,
                    int32_t *image5, oracle_mc_result *res, uint32_t *fate, float *fate_e) {
    const monte_mc_geom *g = S->g;
    const int q = S->o->quirks;
    const double num_p = g->angle0_deg + g->angle_step_deg * view;
    const double Dsd = g->dso + g->dod;
    const double inv_pixel = 1.0 / g->pixel;
    uint64_t steps = 0;
    res->histories++;
    if (R->mode == ORACLE_RNG_PHILOX) rng_history(R, S->o->seed, hid);

    double u_jy = 0.5, u_jz = 0.5, u_e;
    if (R->mode == ORACLE_RNG_MT) {
        if (q & OQ_EXTRA_DRAW) (void)mt_real3(R->mt);                  /* CBCT_real2.cpp:176 */
        if (g->source_mode == MONTE_MC_SOURCE_CONE) { u_jy = mt_real3(R->mt); u_jz = mt_real3(R->mt); }
        u_e = (S->spec && S->spec->n_bins > 0) ? mt_real3(R->mt) : 0.5;  /* CBCT_real325im.cu:492 */
    } else {
        double a, b, c, d;
        rng_pair(R, STREAM_SOURCE, 0, &a, &b);
        rng_pair(R, STREAM_SOURCE, 1, &c, &d);
        if (g->source_mode == MONTE_MC_SOURCE_CONE) { u_jy = a; u_jz = b; }
        u_e = c;
    }
    double E = sample_energy(S, u_e);
    if (q & OQ_ENERGY_INT) E = (int)E;

    /* ---- source / ray set-up, CBCT_real2.cpp:205-266 ---- */
    photon_t P; memset(&P, 0, sizeof(P));
    P.x = -g->dso; P.y = 0; P.z = 0;
    double yl, zl;
    if (q & OQ_PIXEL_OFFSET) { yl = (g->half - 0.5 * g->pixel) - g->pixel * i; zl = (g->half - 0.5 * g->pixel) - g->pixel * j; }
    else { yl = g->half - g->pixel * (i + u_jy); zl = g->half - g->pixel * (j + u_jz); }   /* pixel centre, :476-477 */
    double cos_theta_a, sin_theta_a, cos_phi_a, sin_phi_a;
    double phia = atan(yl / Dsd);
    phia += M_PI * num_p / 180.;
    if (q & OQ_POLAR_ATAN) {
        double theta_a = 0.5 * M_PI - atan(zl / Dsd);
        cos_theta_a = cos(theta_a); sin_theta_a = sin(theta_a);
    } else {                                   /* exact aim at (Dsd, yl, zl) */
        double n = sqrt(Dsd * Dsd + yl * yl + zl * zl);
        cos_theta_a = zl / n; sin_theta_a = sqrt(Dsd * Dsd + yl * yl) / n;
    }
    const int ring = g->detector_shape == MONTE_MC_DETECTOR_RING;
    if (ring) {                                /* from the origin at bin (i, j) of the ring */
        P.x = 0;
        phia = 2 * M_PI * (i + u_jy) / g->ny + M_PI * num_p / 180.;
        double n = sqrt(g->ring_radius * g->ring_radius + zl * zl);
        cos_theta_a = zl / n; sin_theta_a = g->ring_radius / n;
    }
    sin_phi_a = sin(phia); cos_phi_a = cos(phia);
    const double sin_theta_a0 = sin_theta_a, cos_theta_a0 = cos_theta_a, sin_phi_a0 = sin_phi_a, cos_phi_a0 = cos_phi_a;
    double primary_x = P.x * cos(M_PI * num_p / 180) - P.y * sin(M_PI * num_p / 180);
    double primary_y = P.x * sin(M_PI * num_p / 180) + P.y * cos(M_PI * num_p / 180);
    P.x = primary_x; P.y = primary_y;
    int missed = 0;
    if (q & OQ_PREADVANCE) {                   /* CBCT_real2.cpp:246-266 */
        double pv[3];
        pv[0] = g->dod * cos(M_PI * num_p / 180) - yl * sin(M_PI * num_p / 180) - primary_x;
        pv[1] = g->dod * sin(M_PI * num_p / 180) + yl * cos(M_PI * num_p / 180) - primary_y;
        pv[2] = zl;
        double to = (double)(g->dso - S->o->preadvance_start) / Dsd;
        for (int n = 0; n < 3; n++) pv[n] *= to;
        P.x += pv[0]; P.y += pv[1]; P.z += pv[2];
    } else {                                   /* analytic flight to the clip box (slab method) */
        const double d[3] = {sin_theta_a * cos_phi_a, sin_theta_a * sin_phi_a, cos_theta_a};
        const double o3[3] = {P.x, P.y, P.z};
        double t0 = 0, t1 = 1e30;
        for (int a = 0; a < 3; a++) {
            if (d[a] != 0) {
                double ta = (S->vol->clip_lo[a] - o3[a]) / d[a], tb = (S->vol->clip_hi[a] - o3[a]) / d[a];
                if (ta > tb) { double t = ta; ta = tb; tb = t; }
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
            } else if (o3[a] < S->vol->clip_lo[a] || o3[a] >= S->vol->clip_hi[a]) t1 = -1;
        }
        if (t0 >= t1) missed = 1;
        else { P.x += t0 * d[0]; P.y += t0 * d[1]; P.z += t0 * d[2]; }
    }

    double cos_theta_a_new = 1., cos_phi_a_new = 0., sin_theta_a_new = 0., sin_phi_a_new = 1.;   /* :193 */
    int collided = 0;
    P.at_entry = 1;
    if (!missed) collided = delta_sampling(S, R, &P, E, sin_theta_a, cos_theta_a, sin_phi_a, cos_phi_a, &steps);
    P.at_entry = 0;

    const double cr = cos(M_PI * -num_p / 180), sr = sin(M_PI * -num_p / 180);   /* rotate back by -beta */
    uint32_t kind = 0, bin = 0, nint = 0;
    int is_primary;
    if (q & OQ_PREADVANCE) {
        double x_r = P.x * cr - P.y * sr;
        is_primary = (q & OQ_PRIMARY_STRICT) ? (x_r > g->dod) : (x_r >= g->dod);     /* :287 / .cu:567 */
    } else is_primary = !collided;

    if (is_primary) {
        res->primaries++;
        res->sum_e_primary += E;
        kind = 1;
        double x_r, y_r;
        if (missed) { x_r = Dsd - g->dso; y_r = yl; P.z = zl; }
        else { x_r = P.x * cr - P.y * sr; y_r = P.x * sr + P.y * cr; }
        double d_z = ((P.z - 0) / (x_r - (-g->dso))) * g->dod + g->dso * P.z / (x_r + g->dso);   /* .cu:571-572 */
        double d_y = ((y_r - 0) / (x_r - (-g->dso))) * g->dod + g->dso * y_r / (x_r + g->dso);
        int ry = det_bin(d_y, inv_pixel, g->ny), rx = det_bin(d_z, inv_pixel, g->nx);
        if (ring) {                                /* the bin it was aimed at, by the same cylinder intersection as a scattered hit */
            ry = i; rx = j;
            if (!missed && !ring_hit(g, 0, 0, 0, x_r, y_r, P.z, &ry, &rx)) { ry = i; rx = j; }
        }
        if (!(q & OQ_NO_PRIMARY_TALLY) && ry >= 0 && ry < g->ny && rx >= 0 && rx < g->nx) {
            TALLY(image0[ry * g->nx + rx]);
            TALLY(image5[ry * g->nx + rx]);
        }
        bin = (uint32_t)(ry * g->nx + rx);
    } else {
        res->num_scatter++;
        kind = 4;
        int com_flag = 0;
        int escaped = 0;
        for (int a = 0; a < g->max_scatter; a++) {
            /* ---- exit test, CBCT_real325im.cu:613-619 (CBCT_real2.cpp:343-348 has the y bug) ---- */
            if (q & OQ_EXIT_TEST_BUG) {
                double phi0 = 0;
                double x_rotate_c = P.x * cos(M_PI * -phi0 / 180) - P.y * sin(M_PI * -phi0 / 180);
                double y_rotate_c = P.x * cos(M_PI * -phi0 / 180) - P.y * sin(M_PI * -phi0 / 180);
                if (x_rotate_c >= g->dod || fabs(y_rotate_c) >= g->half || fabs(P.z) > g->half) break;
            } else {
                double x_rotate_c = P.x * cr - P.y * sr, y_rotate_c = P.x * sr + P.y * cr;
                if (ring) { if (P.x * P.x + P.y * P.y >= g->ring_radius * g->ring_radius || fabs(P.z) >= g->half) break; }
                else
                if (x_rotate_c >= g->dod || fabs(y_rotate_c) >= g->half || fabs(P.z) >= g->half) break;
            }
            /* ---- material at the site, CBCT_real325im.cu:624-646 ---- */
            int m;
            if (q & OQ_FIRST_MATERIAL) m = 0;
            else {
                m = label_material(S, lookup(S, P.x, P.y, P.z));
                if (m < 0) m = S->tb->n_materials - 1;     /* the reference's final else (air -> last material) */
            }
            const int k = table_index(E);
            const double ab = S->tb->photo[m][k], coh = S->tb->coh[m][k], mu = S->tb->total[m][k];
            double sc_rand, u_phi = 0;
            if (R->mode == ORACLE_RNG_MT) sc_rand = mt_real3(R->mt);
            else rng_pair(R, STREAM_EVENT, R->n_event++, &sc_rand, &u_phi);
            nint++;
            res->interactions++;
            if (sc_rand <= ab / mu) {                       /* photoelectric, :651-655 */
                res->absorbed++;
                kind = 3;
                break;
            } else if (ab / mu < sc_rand && sc_rand <= (ab + coh) / mu) {   /* coherent, :656-695: no deflection */
                res->coherent++;
                res->num_scatter--;                         /* CBCT_real2.cpp:377 */
                if (g->coherent_mode == MONTE_MC_COHERENT_FORMFACTOR) {
                    /* not in the reference: deflect by an angle drawn from the form factor, then carry on exactly
                     * like the reference's coherent event.  Philox mode: one block per rejection round, phi from the
                     * second word of the interaction's own block (as for Compton). */
                    double cos_theta = 1.0;
                    for (;;) {
                        double r_a, r_acc;
                        if (R->mode == ORACLE_RNG_MT) { r_a = mt_real3(R->mt); r_acc = mt_real3(R->mt); }
                        else rng_pair(R, STREAM_EVENT, R->n_event++, &r_a, &r_acc);
                        if (rayleigh_round(S->tb, m, E, r_a, r_acc, &cos_theta)) break;
                    }
                    double sin_theta = sqrt(1. - cos_theta * cos_theta);
                    if (R->mode == ORACLE_RNG_MT) u_phi = mt_real3(R->mt);
                    double phi = u_phi * 2. * M_PI;
                    if (com_flag) {                         /* same update as Compton's, CBCT_real325im.cu:768-780 */
                        sin_theta_a = sin_theta_a_new; cos_theta_a = cos_theta_a_new;
                        sin_phi_a = sin_phi_a_new; cos_phi_a = cos_phi_a_new;
                    }
                    cos_theta_a_new = cos_theta_a * cos_theta - sin_theta_a * sin_theta * cos(phi);
                    if (cos_theta_a_new < -1) cos_theta_a_new = -1;
                    if (cos_theta_a_new > 1) cos_theta_a_new = 1;
                    sin_theta_a_new = sqrt(1. - pow(cos_theta_a_new, 2));
                    if (sin_theta_a_new > 1e-9) {
                        cos_phi_a_new = (cos_theta_a * cos_phi_a * sin_theta * cos(phi) + sin_theta_a * cos_phi_a * cos_theta - sin_phi_a * sin_theta * sin(phi)) / sin_theta_a_new;
                        sin_phi_a_new = (cos_theta_a * sin_phi_a * sin_theta * cos(phi) + sin_theta_a * sin_phi_a * cos_theta + cos_phi_a * sin_theta * sin(phi)) / sin_theta_a_new;
                    } else { cos_phi_a_new = cos_phi_a; sin_phi_a_new = sin_phi_a; }     /* along +-z: azimuth is immaterial */
                    com_flag = 1;
                }
                if (!com_flag) collided = delta_sampling(S, R, &P, E, sin_theta_a0, cos_theta_a0, sin_phi_a0, cos_phi_a0, &steps);
                else collided = delta_sampling(S, R, &P, E, sin_theta_a_new, cos_theta_a_new, sin_phi_a_new, cos_phi_a_new, &steps);
                if (!(q & OQ_DETECT_UNROT)) {               /* GPU form tallies here, CBCT_real325im.cu:672-694 */
                    double x_rot = P.x * cr - P.y * sr, y_rot = P.x * sr + P.y * cr;
                    double xp_rot = P.x_p * cr - P.y_p * sr, yp_rot = P.x_p * sr + P.y_p * cr;
                    double d_z = ((P.z - P.z_p) / (x_rot - xp_rot)) * g->dod + (x_rot * P.z_p - xp_rot * P.z) / (x_rot - xp_rot);
                    double d_y = ((y_rot - yp_rot) / (x_rot - xp_rot)) * g->dod + (x_rot * yp_rot - xp_rot * y_rot) / (x_rot - xp_rot);
                    int ring_ry = 0, ring_rx = 0;
                    if (ring ? ring_hit(g, xp_rot, yp_rot, P.z_p, x_rot, y_rot, P.z, &ring_ry, &ring_rx)
                             : (x_rot >= g->dod && fabs(d_z) <= g->half && fabs(d_y) <= g->half)) {
                        int ry = det_bin(d_y, inv_pixel, g->ny), rx = det_bin(d_z, inv_pixel, g->nx);
                        if (ring) { ry = ring_ry; rx = ring_rx; }
                        if (ry >= 0 && ry < g->ny && rx >= 0 && rx < g->nx) {
                            TALLY(image5[ry * g->nx + rx]);
                            res->scatter_detected++;
                            res->sum_e_scatter += E;
                            kind = 2; bin = (uint32_t)(ry * g->nx + rx);
                        }
                        break;
                    }
                    if (!collided) { escaped = 1; break; }
                }
            } else {                                         /* Compton, :696-844 */
                res->compton++;
                double lambda = 511.0 / E;
                double lambda_d = 0.;
                uint32_t kahn_round = 0;
                for (;;) {                                   /* Kahn, CBCT_real2.cpp:410-434 */
                    double r1, r2, r3, unused;
                    if (R->mode == ORACLE_RNG_MT) { r1 = mt_real3(R->mt); r2 = mt_real3(R->mt); r3 = mt_real3(R->mt); }
                    else {   /* one Philox block per round: r2, r3 = high 23 bits of the two words, r1 = their 2 x 9 low bits */
                        uint32_t o2[2];
                        philox2x32(R->c0, R->c1hi | (STREAM_EVENT << 22) | (R->n_event++ & 0x3FFFFFu), R->key, o2);
                        r2 = u01_23(o2[0]); r3 = u01_23(o2[1]);
                        r1 = ((double)(((o2[0] & 0x1FFu) << 9) | (o2[1] & 0x1FFu)) + 0.5) * (1.0 / 262144.0);
                        unused = 0; (void)unused;
                    }
                    kahn_round++;
                    if (r1 < (lambda + 2.0) / (9.0 * lambda + 2.0)) {
                        double ro = 1.0 + (2.0 / lambda) * r2;
                        if (r3 <= 4.0 * ((1. / ro) - (1. / (ro * ro)))) { lambda_d = ro * lambda; break; }
                    } else {
                        double ro = (lambda + 2.) / (lambda + 2. * (1. - r2));
                        if (r3 <= 0.5 * (pow((lambda - ro * lambda + 1.), 2) + (1. / ro))) { lambda_d = ro * lambda; break; }
                    }
                }
                double cos_theta = (1. - (lambda_d - lambda));
                if (!(q & OQ_FIRST_COMPTON_Z) && cos_theta < -1) cos_theta = -1;     /* .cu:746-747 */
                double sin_theta = sqrt(1. - pow((cos_theta), 2));
                E = 511. / lambda_d;
                if (q & OQ_ENERGY_INT) E = (int)E;
                double phi;
                if (R->mode == ORACLE_RNG_MT) u_phi = mt_real3(R->mt);
                phi = (q & OQ_PHI_NEG) ? -u_phi * 2. * M_PI : u_phi * 2. * M_PI;
                /* ---- direction update, CBCT_real2.cpp:460-470 / CBCT_real325im.cu:768-780 ---- */
                if ((q & OQ_FIRST_COMPTON_Z) || com_flag) {
                    sin_theta_a = sin_theta_a_new; cos_theta_a = cos_theta_a_new;
                    sin_phi_a = sin_phi_a_new; cos_phi_a = cos_phi_a_new;
                }
                cos_theta_a_new = cos_theta_a * cos_theta - sin_theta_a * sin_theta * cos(phi);
                if (!(q & OQ_FIRST_COMPTON_Z) && cos_theta_a_new < -1) cos_theta_a_new = -1;
                sin_theta_a_new = sqrt(1. - pow(cos_theta_a_new, 2));
                cos_phi_a_new = (cos_theta_a * cos_phi_a * sin_theta * cos(phi) + sin_theta_a * cos_phi_a * cos_theta - sin_phi_a * sin_theta * sin(phi)) / sin_theta_a_new;
                sin_phi_a_new = (cos_theta_a * sin_phi_a * sin_theta * cos(phi) + sin_theta_a * sin_phi_a * cos_theta + cos_phi_a * sin_theta * sin(phi)) / sin_theta_a_new;
                com_flag = 1;
                collided = delta_sampling(S, R, &P, E, sin_theta_a_new, cos_theta_a_new, sin_phi_a_new, cos_phi_a_new, &steps);
                /* ---- scatter detection ---- */
                if (q & OQ_DETECT_UNROT) {                   /* CBCT_real2.cpp:528-583 */
                    double phi_a_result = num_p;
                    double x_rotate = P.x * cos(-phi_a_result) - P.y * sin(-phi_a_result);
                    double d_z = (g->dod - P.x_p) * (P.z - P.z_p) / (P.x - P.x_p) + P.z_p;
                    double d_y = (g->dod - P.x_p) * (P.y - P.y_p) / (P.x - P.x_p) + P.y_p;
                    if (x_rotate >= g->dod && fabs(d_z) < g->half && fabs(d_y) < g->half) {
                        int ry = det_bin(d_y, inv_pixel, g->ny), rx = det_bin(d_z, inv_pixel, g->nx);
                        res->scatter_detected++;
                        res->sum_e_scatter += E;
                        TALLY(image5[ry * g->nx + rx]);
                        kind = 2; bin = (uint32_t)(ry * g->nx + rx);
                    } else if (P.x > g->dod || fabs(P.y) > g->half || fabs(P.z) > g->half) res->num_nd++;
                } else {                                     /* CBCT_real325im.cu:823-843 */
                    double x_rotate = P.x * cr - P.y * sr, y_rotate = P.x * sr + P.y * cr;
                    double x_p_rotate = P.x_p * cr - P.y_p * sr, y_p_rotate = P.x_p * sr + P.y_p * cr;
                    double d_z = ((P.z - P.z_p) / (x_rotate - x_p_rotate)) * g->dod + (x_rotate * P.z_p - x_p_rotate * P.z) / (x_rotate - x_p_rotate);
                    double d_y = ((y_rotate - y_p_rotate) / (x_rotate - x_p_rotate)) * g->dod + (x_rotate * y_p_rotate - x_p_rotate * y_rotate) / (x_rotate - x_p_rotate);
                    int ring_ry = 0, ring_rx = 0;
                    if (ring ? ring_hit(g, x_p_rotate, y_p_rotate, P.z_p, x_rotate, y_rotate, P.z, &ring_ry, &ring_rx)
                             : (x_rotate >= g->dod && fabs(d_z) <= g->half && fabs(d_y) <= g->half)) {
                        int ry = det_bin(d_y, inv_pixel, g->ny), rx = det_bin(d_z, inv_pixel, g->nx);
                        if (ring) { ry = ring_ry; rx = ring_rx; }
                        if (ry >= 0 && ry < g->ny && rx >= 0 && rx < g->nx) {
                            TALLY(image5[ry * g->nx + rx]);
                            res->scatter_detected++;
                            res->sum_e_scatter += E;
                            kind = 2; bin = (uint32_t)(ry * g->nx + rx);
                        }
                        break;
                    }
                    if (!collided) { escaped = 1; break; }
                }
            }
            if (a == g->max_scatter - 1 && kind == 4 && collided) kind = 5;   /* budget exhausted at a collision */
        }
        (void)escaped;
        if (g->max_scatter == 0 && collided) kind = 5;    /* no interaction allowed at all: the budget ends at the first collision */
    }
    res->woodcock_steps += steps;
    if (fate) { *fate = kind | (bin << 8) | (nint << 28); if (fate_e) *fate_e = (float)E; }
}

/* ---------------------------------------------------------------- entry points */
/* Run photons n in [n_begin, n_end) of pixels [i_begin,i_end) x [j_begin,j_end) of views
 * [view_begin, view_end).  image0/image5: [n_views][ny][nx] int32, accumulated into.
 * MT mode is sequential in the reference's loop order (view, i, j, photon) when n_threads == 1;
 * with more threads each (view, detector row) task gets its own MT stream (statistical use only).
 * fates (nullable): one record per history of a single view, index (i*nx + j)*per + n.          */
int oracle_mc_run(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                  const oracle_mc_tables *tb, const monte_mc_spectrum *spec, const oracle_mc_opts *o,
                  uint32_t per, uint32_t n_begin, uint32_t n_end, int view_begin, int view_end,
                  int i_begin, int i_end, int j_begin, int j_end,
                  int32_t *image0, int32_t *image5, oracle_mc_result *result, uint32_t *fates, float *fate_e) {
    scene_t S = {g, vol, labels, tb, spec, o, 0xffffffffu};
    if (vol->majorant_mode == MONTE_MC_MAJORANT_PRESENT) {
        S.present = 0;
        const size_t nvox = (size_t)vol->nx * vol->ny * vol->nz;
        for (size_t i = 0; i < nvox; i++) {
            const int m = label_material(&S, labels[i]);
            if (m >= 0) S.present |= 1u << m;
        }
    }
    memset(result, 0, sizeof(*result));
    const size_t npix = (size_t)g->ny * g->nx;
    int nthreads = o->n_threads;
    if (o->rng_mode == ORACLE_RNG_MT && nthreads != 1 && (o->quirks & OQ_EXTRA_DRAW)) nthreads = 1;
    mt_state shared_mt;
    mt_seed(&shared_mt, (uint32_t)o->seed);
    const int nrows = i_end - i_begin;
    const long n_tasks = (long)(view_end - view_begin) * nrows;     /* one task = one detector row of one view */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads()) if (nthreads != 1)
    for (long task = 0; task < n_tasks; task++) {
        const int view = view_begin + (int)(task / nrows), i = i_begin + (int)(task % nrows);
        oracle_mc_result local; memset(&local, 0, sizeof(local));
        mt_state my_mt;
        rng_t R; memset(&R, 0, sizeof(R));
        R.mode = o->rng_mode;
        if (nthreads == 1) R.mt = &shared_mt;
        else { mt_seed(&my_mt, (uint32_t)(o->seed + 7919u * (uint32_t)(view * g->ny + i))); R.mt = &my_mt; }
        int32_t *im0 = image0 + (size_t)view * npix, *im5 = image5 + (size_t)view * npix;
        for (int j = j_begin; j < j_end; j++)
            for (uint32_t n = n_begin; n < n_end; n++) {
                const size_t pix = (size_t)i * g->nx + j;
                const uint64_t hid = ((uint64_t)view * npix + pix) * per + n;
                uint32_t *f = fates ? fates + (pix * per + n) : NULL;
                float *fe = (fates && fate_e) ? fate_e + (pix * per + n) : NULL;
                history(&S, &R, view, i, j, hid, im0, im5, &local, f, fe);
            }
#pragma omp critical
        {
            result->histories += local.histories; result->primaries += local.primaries;
            result->scatter_detected += local.scatter_detected; result->absorbed += local.absorbed;
            result->interactions += local.interactions; result->coherent += local.coherent;
            result->compton += local.compton; result->woodcock_steps += local.woodcock_steps;
            result->num_scatter += local.num_scatter; result->num_nd += local.num_nd;
            result->sum_e_primary += local.sum_e_primary; result->sum_e_scatter += local.sum_e_scatter;
        }
    }
    return 0;
}

/* Deterministic primary projection: line integral of mu(E) from the source to the pixel centre,
 * exact voxel traversal in double.  map[view][i][j].  (The role monte_cpp/projection.cpp:74-125
 * plays in 2-D; the variance-free limit of -ln(image0/per).)                                    */
int oracle_project_primary(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                           const oracle_mc_tables *tb, double keV, int view_begin, int view_end, float *map) {
    const int k = table_index(keV);
    const double Dsd = g->dso + g->dod;
    double mu[256];
    for (int l = 0; l < 256; l++) {
        int m = l == 0 ? -1 : (l <= tb->n_materials ? l - 1 : tb->n_materials - 1);
        mu[l] = m < 0 ? 0.0 : tb->total[m][k] * tb->density[m];
    }
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int view = view_begin; view < view_end; view++)
        for (int i = 0; i < g->ny; i++)
            for (int j = 0; j < g->nx; j++) {
                const double beta = M_PI * (g->angle0_deg + g->angle_step_deg * view) / 180;
                const double cb = cos(beta), sb = sin(beta);
                const double yl = g->half - g->pixel * (i + 0.5), zl = g->half - g->pixel * (j + 0.5);
                const double src[3] = {-g->dso * cb, -g->dso * sb, 0};
                const double n = sqrt(Dsd * Dsd + yl * yl + zl * zl);
                const double d[3] = {(Dsd * cb - yl * sb) / n, (Dsd * sb + yl * cb) / n, zl / n};
                double t0 = 0, t1 = 1e30;
                for (int a = 0; a < 3; a++) {
                    if (d[a] != 0) {
                        double ta = (vol->clip_lo[a] - src[a]) / d[a], tb2 = (vol->clip_hi[a] - src[a]) / d[a];
                        if (ta > tb2) { double t = ta; ta = tb2; tb2 = t; }
                        if (ta > t0) t0 = ta;
                        if (tb2 < t1) t1 = tb2;
                    } else if (src[a] < vol->clip_lo[a] || src[a] >= vol->clip_hi[a]) t1 = -1;
                }
                double acc = 0;
                if (t0 < t1) {
                    /* Amanatides-Woo traversal of the voxel grid between t0 and t1 */
                    const double inv = 1.0 / vol->pitch;
                    double t = t0;
                    int idx[3], step[3];
                    double tnext[3], dt[3];
                    const int dims[3] = {vol->nx, vol->ny, vol->nz};
                    const double eps = 1e-9;
                    for (int a = 0; a < 3; a++) {
                        double p = src[a] + (t0 + eps) * d[a];
                        idx[a] = (int)floor((p - vol->origin[a]) * inv);
                        step[a] = d[a] > 0 ? 1 : -1;
                        if (d[a] != 0) {
                            double edge = vol->origin[a] + (idx[a] + (d[a] > 0 ? 1 : 0)) * vol->pitch;
                            tnext[a] = (edge - src[a]) / d[a];
                            dt[a] = vol->pitch / fabs(d[a]);
                        } else { tnext[a] = 1e30; dt[a] = 1e30; }
                    }
                    while (t < t1) {
                        int a = tnext[0] <= tnext[1] ? (tnext[0] <= tnext[2] ? 0 : 2) : (tnext[1] <= tnext[2] ? 1 : 2);
                        double te = tnext[a] < t1 ? tnext[a] : t1;
                        if (idx[0] >= 0 && idx[1] >= 0 && idx[2] >= 0 && idx[0] < dims[0] && idx[1] < dims[1] && idx[2] < dims[2]) {
                            int l = labels[((size_t)idx[2] * vol->ny + idx[1]) * vol->nx + idx[0]];
                            if (te > t) acc += mu[l] * (te - t);
                        }
                        t = te;
                        idx[a] += step[a];
                        tnext[a] += dt[a];
                    }
                }
                map[((size_t)view * g->ny + i) * g->nx + j] = (float)acc;
            }
    return 0;
}

/* counts -> line-integral map, CBCT_real325im.cu:267-285 */
void oracle_counts_to_map(const int32_t *counts, size_t n, int32_t per, float *map) {
    for (size_t i = 0; i < n; i++) {
        int c = counts[i];
        if (c > per) c = per;
        if (c == 0) c = 1;
        map[i] = -log((double)c) + (double)logf((float)per);   /* C++ overloads: log(int)->double, log(float)->float */
    }
}

/* OpenMP team size of every oracle entry point (the FDK loops have no per-call thread count): n > 0 sets it, the
 * return value is the size in force.  bench.py's CPU legs call it with the host's core count because a launcher may
 * export OMP_NUM_THREADS=1 (torch.distributed.run does), and report the number returned. */
int oracle_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
