// mc.cu — photon-history Monte Carlo transport for sm_100a and its C-ABI entry points.
//
// Replaces the `projection` kernel of the reference (monte_cu/CBCT_real325im.cu:425-860; CPU form
// monte_cpp/CBCT_real2.cpp:169-595) — same physics, different machine mapping:
//   reference: one thread per detector PIXEL runs views x photons serially, cuRAND MRG32k3a state
//              per pixel, Woodcock steps through ~200 cm of air, 16 table pointers in global memory,
//              one global atomicAdd per detected photon.
//   here:      one lane per HISTORY.  Warps pull units of consecutive histories from a global
//              counter and refill finished lanes from the unit (no lane waits for the slowest
//              history of its warp).  Counter-based Philox2x32-10 keyed by (seed, history id):
//              results do not depend on the partition across lanes, CTAs or GPUs.  Air outside the
//              clip box is crossed analytically (identical distribution: every tentative collision
//              in air is rejected, CBCT_real2.cpp:780-785).  Cross-section tables and the majorant
//              per keV live in shared memory.  Unscattered photons land in the pixel they were aimed
//              at, so they are counted in a register and flushed with one atomic per (lane, pixel).
// The oracle (oracle/mc_oracle.c, quirks=0, Philox mode) consumes the same variates in the same
// order; tests compare the two history by history.
#include "common.cuh"
#include "nccl_dl.cuh"
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <thread>

namespace monte {

constexpr int MC_THREADS = 256;
#ifndef MONTE_MC_UNIT
#define MONTE_MC_UNIT 1024
#endif
constexpr int MC_UNIT = MONTE_MC_UNIT;   // histories per work unit (one warp)
constexpr int TAB_ROWS = MONTE_MC_TABLE_ROWS;

struct McSceneDev {
    const uint8_t *labels;
    int nx, ny, nz;
    float inv_pitch;
    float org[3], clip_lo[3], clip_hi[3];
    float vox_off[3];           // -org * inv_pitch - 0.5 (one rounding): voxel index = round-to-nearest(pos * inv_pitch + vox_off)
    const float4 *tab;          // [n_mat][201]: {mu_m/mu_max, photo/total, (photo+coh)/total, 0}
    const float *inv_mumax;     // [201]
    int n_mat;
    const float *cdf;           // [n_bins+1] or null
    int n_bins;
    float bin_keV, mono_keV;
    const float2 *view_cs;      // [n_views] cos, sin of the view angle
    int n_views, det_ny, det_nx;
    float pixel, inv_pixel, half, dso, dod, dsd;
    int source_mode, max_scatter;
    int eid;                    // energy-integrating detector: tally (int)(E*16+0.5) instead of 1
    int ring;                   // detector_shape RING: source at the origin, cylinder of radius ring_r about the z axis
    float ring_r, ring_r2, ring_dphi, ring_kphi;   // radius, its square, 2 pi / ny, ny / 2 pi
};

struct PhiloxKeys { uint32_t rk[10]; };   // key + r * 0x9E3779B9, r = 0..9 (see philox2x32_10)

struct McLaunch {
    McSceneDev sc;
    PhiloxKeys key;               // the ten Philox round keys of this launch's seed
    int view_begin;
    uint32_t n_begin, cnt, per;   // photons [n_begin, n_begin+cnt) of every pixel; per = id-space size
    unsigned long long total;     // histories of this launch
    unsigned long long n_units;
    int32_t *image0, *image5;
    unsigned long long *stats;    // MONTE_MC_STATS_WORDS accumulators (nullable)
    unsigned long long *work;     // unit counter (zeroed by the host)
    uint32_t *fates;              // RECORD only
    float *fate_e;
    uint32_t off_inv, off_cdf, off_ray, off_invlo, off_slots;   // shared-memory layout, byte offsets (see the kernel)
    uint32_t vote_bias;           // (128 - T) in every byte: a phase runs on a vote when >= T lanes wait for it (T = 14)
    uint32_t collide_check;       // 0: the clip box lies inside the detector-side bounds of :613-619 for every view, so no
                                  // collision site can fail that test and COLLIDE skips it (host: launch_mc)
    uint32_t n_vox_m1;            // voxels of the label volume - 1: the one clamp of a label address
    // RAYLEIGH instantiations only (appended: the parameter offsets of everything above do not move)
    const float *ray;             // [n_mat][2][ray_n]: x^2 grid, then cumulative F^2 (monte_mc_xs.ff_x2 / ff_cum)
    int ray_n;
    // CLEAR instantiations only (monte_mc_volume.tracking_mode == MONTE_MC_TRACK_CLEARANCE)
    const uint8_t *clear;         // clearance grid [cgz][cgy][cgx] (monte_mc_clearance_grid, clipped to 127)
    const float *inv_mulo;        // [201] 1 / majorant of every material but the heavy one
    int cgx, cgy, cshift;         // grid dims and log2 of the cell side in voxels
    float cunit;                  // cm per grid unit (half a cell side)
    const float *clear_thr;       // [201] clearance (cm) above which the light majorant is used: 0 (CLEARANCE) or the
                                  // break-even distance -ln(1 - mu_light/mu_max)/mu_light (ADAPTIVE)
    unsigned coct;                // DIRECTIONAL: cells per grid (eight grids, one per direction octant); 0: one grid
};

// stats word indices
enum { ST_HIST = 0, ST_PRIM, ST_SCAT, ST_ABS, ST_INT, ST_COH, ST_COMP, ST_STEPS, ST_EPRIM, ST_ESCAT };

// ---- Philox2x32-10 (Salmon, Moraes, Dror, Shaw, SC'11): counter (c0,c1), 32-bit key ---------
// rk[r] = key + r * 0x9E3779B9 (the key schedule) is the same for every history of a launch: the host computes the ten
// round keys once and the kernel reads them from the parameter (constant) bank as an operand of the round's LOP3,
// instead of every lane bumping its own copy every round (3.1 % of the kernel's instructions, ncu source page r02).
__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, const PhiloxKeys &k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi = __umulhi(0xD256D193u, c0);
        const uint32_t lo = 0xD256D193u * c0;
        c0 = hi ^ k.rk[r] ^ c1;
        c1 = lo;
    }
    return make_uint2(c0, c1);
}
// 23-bit uniform in (0,1): ((x>>9)+0.5)*2^-23, exact in fp32, never 0 or 1.  Built from the bit
// pattern of 1.m (no I2F on the XU pipe): (1 + k*2^-23) - (1 - 2^-24) = (k + 0.5)*2^-23 exactly.
__device__ __forceinline__ float u01(uint32_t x) {
    return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994f;
}

#define STREAM_FLIGHT 0u
#define STREAM_EVENT  1u
#define STREAM_SOURCE 2u

// ------------------------------------------------------------------------------------------------
// The transport loop of one history is a sequence of phases of very different cost and frequency
// (Woodcock step / collision bookkeeping / one Kahn round / end-of-history tally + source set-up).
// Lanes of a warp are in different phases at any time.  v1 let every lane run its own control flow
// (8.9 of 32 lanes active per issued instruction); v2 let the warp vote and execute only the phase
// most lanes wait for (16.9 lanes), but a lane that owns exactly one history is often in a minority
// phase.  This kernel (v3) parks K histories per lane in shared memory: every lane owns K histories
// ("slots"); their state lives in shared memory (struct-of-arrays, column = lane, so accesses are
// conflict-free) and only a 4-bit one-hot phase per slot stays in a register.  Each iteration the
// warp votes for the phase in which most LANES have at least one slot, every such lane loads that
// slot, advances it by one phase and stores it back.  With K = 4 nearly every lane has work in the
// winning phase.  Units are chained without draining the warp.  Per-history variates are unchanged
// (counter-based), so results are bit-identical to v1/v2.
// ------------------------------------------------------------------------------------------------
enum : uint32_t { P_REFILL = 1u, P_STEP = 2u, P_COLLIDE = 4u, P_COMPTON = 8u };
// Slot state = three (RECORD: four) 16-byte groups per slot and lane, laid out [group][slot][lane]
// so that a warp's LDS.128 / STS.128 touches 512 contiguous bytes (conflict-free):
//   G_POS : x, y, z, E            G_DIR : dx, dy, dz, u_phi
//   G_ID  : c0, META, CTR, PIXVIEW            G_REC : record index (fate dump only)
// META: bits 0-7 kE, 8-11 nint, 12-14 material, 15 pending-detect, 16 coherent event waiting for its angle
//       (RAYLEIGH kernels), 17-23 clearance of the cell the photon is in, in grid units (CLEAR kernels),
//       24-31 high byte of the history id
// CTR : bits 0-19 flight-stream index, 20-31 event-stream index
// PIXVIEW: bits 0-19 pixel, 20-31 view
enum { G_POS = 0, G_DIR = 1, G_ID = 2, G_REC = 3 };
constexpr int mc_slot_groups(bool record) { return record ? 4 : 3; }

// y(v) by linear interpolation in a table (xs ascending, n >= 2), v clamped to the table's range
__device__ __forceinline__ float ray_interp(const float *xs, const float *ys, int n, float v) {
    int lo = 0, hi = n - 2;                      // largest i <= n-2 with xs[i] <= v
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (xs[mid] <= v) lo = mid; else hi = mid - 1; }
    const float x0 = xs[lo], w = xs[lo + 1] - x0;
    const float t = w > 0.f ? fminf(fmaxf(__fdividef(v - x0, w), 0.f), 1.f) : 0.f;
    return fmaf(t, ys[lo + 1] - ys[lo], ys[lo]);
}

// RAYLEIGH = true: monte_mc_geom.coherent_mode == MONTE_MC_COHERENT_FORMFACTOR (off in parity mode).
// A coherent event then waits in the COMPTON phase for its angle: one rejection round per visit, like Kahn's.
// CLEAR = true: monte_mc_volume.tracking_mode == MONTE_MC_TRACK_CLEARANCE, the two-level majorant.  Every slot
// carries the clearance of its current cell (META bits 17-23, fetched together with the label at the end of a
// step); a step that starts with clearance > 0 is sampled with the majorant of the lighter materials and cut
// at the clearance radius.  Same collision-site distribution as the reference's loop, far fewer virtual
// collisions when a dense insert sets the global majorant (C4: 23 -> ~5 steps per history).
// RING = true: monte_mc_geom.detector_shape == MONTE_MC_DETECTOR_RING: source at the origin, detector bins on a
// cylinder about the z axis.  Its own instantiations (with and without RAYLEIGH / CLEAR): the flat-panel kernels carry none of it.
template <bool RECORD, int K, int MINB = 3, int NSTEP = 2, bool RAYLEIGH = false, bool CLEAR = false, bool RING = false>
__global__ void __launch_bounds__(MC_THREADS, MINB)
mc_transport_kernel_v3(const __grid_constant__ McLaunch P) {
    MONTE_DYN_SMEM(float4, s_mem);
    const McSceneDev &sc = P.sc;
    // layout (byte offsets computed once on the host, McLaunch::off_*: with a run-time material count the compiler
    // re-derived these pointers inside the loop): tables [n_mat*201] float4 | 1/mu_max [201 (+3)] | spectrum CDF
    // [n_bins+1, padded to 4] | RAYLEIGH: [n_mat][2][ray_n padded to 4] | CLEAR: 1/mu_light [201 (+3)], thresholds
    // [201 (+3)] | slots | per-warp source-ray cache
    constexpr int NG = mc_slot_groups(RECORD);
    constexpr int GSTRIDE = K * 32;                                    // uint4 per group per warp
    char *s_base = reinterpret_cast<char *>(s_mem);
    float4 *s_tab = s_mem;
    float *s_inv = reinterpret_cast<float *>(s_base + P.off_inv);
    float *s_cdf = reinterpret_cast<float *>(s_base + P.off_cdf);
    float *s_ray = reinterpret_cast<float *>(s_base + P.off_ray);
    const int ray_stride = RAYLEIGH ? ((P.ray_n + 3) & ~3) : 0;
    float *s_invlo = reinterpret_cast<float *>(s_base + P.off_invlo);
    float *s_thr = s_invlo + TAB_ROWS + 3;
    uint4 *s_slots = reinterpret_cast<uint4 *>(s_base + P.off_slots) + (threadIdx.x >> 5) * (NG * GSTRIDE);
    // per-warp cache of the last pixel's source ray (pencil source, one energy: the `per` photons of a pixel share it)
    uint4 *s_src = reinterpret_cast<uint4 *>(s_base + P.off_slots) + (MC_THREADS / 32) * (NG * GSTRIDE) + (threadIdx.x >> 5) * 3;
    for (int i = threadIdx.x; i < sc.n_mat * TAB_ROWS; i += MC_THREADS) s_tab[i] = sc.tab[i];
    for (int i = threadIdx.x; i < TAB_ROWS; i += MC_THREADS) s_inv[i] = sc.inv_mumax[i];
    for (int i = threadIdx.x; i <= sc.n_bins && sc.n_bins > 0; i += MC_THREADS) s_cdf[i] = sc.cdf[i];
    if (RAYLEIGH)
        for (int i = threadIdx.x; i < sc.n_mat * 2 * P.ray_n; i += MC_THREADS)
            s_ray[(i / P.ray_n) * ray_stride + i % P.ray_n] = P.ray[i];
    if (CLEAR)
        for (int i = threadIdx.x; i < TAB_ROWS; i += MC_THREADS) { s_invlo[i] = P.inv_mulo[i]; s_thr[i] = P.clear_thr[i]; }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t npix = (uint32_t)(sc.det_ny * sc.det_nx);
    const float *vox_off = sc.vox_off;                                 // (parameter bank: an operand, not registers)
    constexpr uint32_t ALL = (K == 6) ? 0x111111u : (K == 5) ? 0x11111u : (K == 4) ? 0x1111u : (K == 3) ? 0x0111u : (K == 2) ? 0x0011u : 0x0001u;

#define GRP(g) (slot[(g) * GSTRIDE])
#define WORD(g, w) (reinterpret_cast<uint32_t *>(slot + (g) * GSTRIDE)[w])
#define WORDF(g, w) (reinterpret_cast<float *>(slot + (g) * GSTRIDE)[w])

    const bool src_cached = sc.source_mode != MONTE_MC_SOURCE_CONE && sc.n_bins == 0;
    if ((threadIdx.x & 31) == 0) s_src[2] = make_uint4(0xffffffffu, 0u, 0u, 0u);   // tag: no pixel yet
    __syncwarp();
    uint32_t st = ALL * P_REFILL;                                      // one-hot phase per slot
#pragma unroll
    for (int j = 0; j < K; j++) s_slots[G_ID * GSTRIDE + j * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);   // no pending detection
    uint32_t cur_pv = 0xffffffffu, prim_cnt = 0;
    uint32_t c_hist = 0, c_prim = 0, c_scat = 0, c_abs = 0, c_int = 0, c_coh = 0, c_comp = 0, c_steps = 0;
    unsigned long long e_prim = 0, e_scat = 0;                         // fixed point, 1/1024 keV
    // current unit (warp-uniform)
    uint32_t unit_cnt = 0, next_off = 0, n0 = 0, v0 = 0, p0 = 0, pi0 = 0, pj0 = 0;
    bool grid_done = false;
    const bool few_pixels_per_unit = P.cnt >= (uint32_t)MC_UNIT / 4u;  // a unit then spans at most six pixels

    for (;;) {
        // ---------------- vote: the phase in which most lanes have a slot waiting ---------------
        uint32_t onehot = 0;
        if (st & (ALL * P_REFILL)) onehot |= 1u;
        if (st & (ALL * P_STEP)) onehot |= 1u << 8;
        if (st & (ALL * P_COLLIDE)) onehot |= 1u << 16;
        if (st & (ALL * P_COMPTON)) onehot |= 1u << 24;
        const uint32_t cnts = __reduce_add_sync(0xffffffffu, onehot);
        if (cnts == 0u) break;                                          // every slot of every lane is done
        // Every phase in which at least `second_min` (14) lanes wait is run on this one vote, in pipeline order STEP ->
        // COLLIDE -> COMPTON -> REFILL (each feeds the next, so the stale counts of the later ones can only have grown).
        // Bit 7 of byte p of `todo` = phase p qualifies: count + (128 - T) carries into bit 7 iff count >= T (counts <= 32,
        // no carry between bytes).  Near the end of the grid no phase qualifies: then the fullest one runs alone.
        uint32_t todo = (cnts + P.vote_bias) & 0x80808080u;
        if (todo == 0u) {
            const uint32_t c_r = cnts & 0xFFu, c_s = (cnts >> 8) & 0xFFu, c_c = (cnts >> 16) & 0xFFu, c_k = cnts >> 24;
            const uint32_t mx = max(max(c_r, c_s), max(c_c, c_k));
            todo = c_s == mx ? 0x8000u : c_c == mx ? 0x800000u : c_k == mx ? 0x80000000u : 0x80u;
        }
        // the qualifying phases, straight-line (no phase variable, no compare chain)
#define MC_PHASE_SLOT(PH)                                                                                   \
        const uint32_t mine = st & (ALL * (PH));                                                              \
        const bool active = mine != 0u;                                                                       \
        const int j = active ? ((__ffs(mine) - 1) >> 2) : 0;            /* my slot in that phase */           \
        uint4 *slot = s_slots + j * 32 + lane;                                                                \
        const uint32_t clr = ~(0xFu << (4 * j));
        if (todo & 0x8000u) do {                                        // ======== STEP
        MC_PHASE_SLOT(P_STEP)
            if (!active) break;
            // ---------------- Woodcock steps, CBCT_real325im.cu:886-968 ---------------
            // Two slots of the lane are advanced per visit when it has two waiting (the second one is
            // predicated off otherwise): the two Philox chains and the two label fetches are independent,
            // so they overlap instead of each stalling the warp on its own fixed-latency chain ("wait" was
            // the top stall reason), and the vote / dispatch overhead per step drops.  Measured at C2:
            // one slot 8.52 ms, two 8.25 ms, three 9.05 ms per 1e8 histories.
            constexpr int NS = NSTEP;                              // slots advanced per visit
            bool en2[NS]; int jj[NS]; uint4 *sl[NS];
            {
                uint32_t rest = mine;
#pragma unroll
                for (int q = 0; q < NS; q++) {
                    en2[q] = rest != 0u;
                    jj[q] = rest ? ((__ffs(rest) - 1) >> 2) : j;
                    sl[q] = s_slots + jj[q] * 32 + lane;
                    rest &= rest - 1u;
                }
            }
            uint4 id2[NS]; float4 pos2[NS]; uint2 r2[NS]; bool inside2[NS]; int lab2[NS];
            bool lo2[NS]; uint32_t qn2[NS];                        // CLEAR: sampled with the light majorant; clearance at the new site
#pragma unroll
            for (int q = 0; q < NS; q++) {
                id2[q] = sl[q][G_ID * GSTRIDE];
                pos2[q] = *reinterpret_cast<float4 *>(&sl[q][G_POS * GSTRIDE]);
            }
#pragma unroll
            for (int q = 0; q < NS; q++)
                r2[q] = philox2x32_10(id2[q].x, (id2[q].y & 0xFF000000u) | (STREAM_FLIGHT << 22) | (id2[q].z & 0xFFFFFu), P.key);
#pragma unroll
            for (int q = 0; q < NS; q++) {
                const float4 dir = *reinterpret_cast<float4 *>(&sl[q][G_DIR * GSTRIDE]);
                float sp;
                bool cut = false;
                lo2[q] = false; qn2[q] = 0u;
                if (CLEAR) {
                    const uint32_t qc = (id2[q].y >> 17) & 0x7Fu;  // clearance of the cell the step starts in
                    const float dcl = (float)qc * P.cunit;
                    lo2[q] = dcl > s_thr[id2[q].y & 0xFF];         // CLEARANCE: threshold 0, i.e. wherever there is clearance
                    sp = -__logf(u01(r2[q].x)) * (lo2[q] ? s_invlo[id2[q].y & 0xFF] : s_inv[id2[q].y & 0xFF]);
                    cut = lo2[q] && sp > dcl;                      // would leave the cleared ball: stop at its surface, no collision
                    sp = cut ? dcl : sp;
                } else sp = -__logf(u01(r2[q].x)) * s_inv[id2[q].y & 0xFF];
                pos2[q].x = fmaf(sp, dir.x, pos2[q].x); pos2[q].y = fmaf(sp, dir.y, pos2[q].y); pos2[q].z = fmaf(sp, dir.z, pos2[q].z);
                inside2[q] = pos2[q].x >= sc.clip_lo[0] && pos2[q].x < sc.clip_hi[0] && pos2[q].y >= sc.clip_lo[1] && pos2[q].y < sc.clip_hi[1] &&
                             pos2[q].z >= sc.clip_lo[2] && pos2[q].z < sc.clip_hi[2];
                int ix = __float_as_int(fmaf(pos2[q].x, sc.inv_pitch, vox_off[0]) + 12582912.0f) - 0x4B400000;
                int iy = __float_as_int(fmaf(pos2[q].y, sc.inv_pitch, vox_off[1]) + 12582912.0f) - 0x4B400000;
                int iz = __float_as_int(fmaf(pos2[q].z, sc.inv_pitch, vox_off[2]) + 12582912.0f) - 0x4B400000;
                // the clip box lies inside the volume (scene_upload checks it), so a position that passed the clip test
                // indexes a voxel; a coordinate exactly on an outer face may round one voxel out: safety only, one clamp
                // of the linear address (CLEAR kernels: per axis, the cell address below needs the three indices)
                if (CLEAR) {
                    ix = (int)min((unsigned)ix, (unsigned)(sc.nx - 1)); iy = (int)min((unsigned)iy, (unsigned)(sc.ny - 1));
                    iz = (int)min((unsigned)iz, (unsigned)(sc.nz - 1));
                }
                lab2[q] = (en2[q] && inside2[q] && !cut) ? __ldg(sc.labels + min((unsigned)(iz * sc.ny + iy) * (unsigned)sc.nx + (unsigned)ix, P.n_vox_m1)) : 0;
                if (CLEAR && en2[q] && inside2[q]) {
                    const unsigned oc = (dir.x > 0.f ? 1u : 0u) | (dir.y > 0.f ? 2u : 0u) | (dir.z > 0.f ? 4u : 0u);   // coct == 0: one grid
                    qn2[q] = __ldg(P.clear + oc * P.coct + ((unsigned)((iz >> P.cshift) * P.cgy + (iy >> P.cshift)) * (unsigned)P.cgx + (unsigned)(ix >> P.cshift)));
                }
            }
#pragma unroll
            for (int q = 0; q < NS; q++) {
                // Straight-line: everything is computed for both slots, only the stores hang on en2[q] (a lane with one waiting
                // slot) -- the lanes of a warp take every outcome on every visit, so a branch saves nothing and costs its
                // resolution (`branch_resolving` was a top-3 stall).  Bitwise &: no short-circuit branches around the table load.
                uint4 *slot = sl[q];                               // GRP/WORD below refer to this slot
                uint32_t meta = id2[q].y;
                const uint32_t ctr = id2[q].z;
                if (CLEAR) meta = (meta & ~0xFE0000u) | (qn2[q] << 17);
                const uint32_t clrq = ~(0xFu << (4 * jj[q]));
                const int kE = meta & 0xFF;
                // Three outcomes:
                //  left the volume -- only air ahead: the history is finished with the REFILL phase (primary tally or scatter
                //    detection; there nearly every lane of a visit has one to finish, here it would be the 3-5 that leave);
                //  accepted tentative collision -- records its material and moves to COLLIDE;
                //  air or rejected (virtual) collision, :941-961 -- nothing more.
                const bool out = !inside2[q];
                const int mat = max(min(lab2[q], sc.n_mat) - 1, 0);
                const float4 tb = s_tab[mat * TAB_ROWS + kE];
                const float ratio = CLEAR && lo2[q] ? tb.w : tb.x;    // acceptance against the majorant the step was sampled with
                const bool accept = (!out) & (lab2[q] != 0) & !(u01(r2[q].y) > ratio);
                const bool moved = (out | accept) & en2[q];
                const uint32_t new_meta = out ? (meta | 0x8000u) : (accept ? ((meta & ~0x7000u) | ((uint32_t)mat << 12)) : meta);
                if (en2[q]) {
                    WORD(G_ID, 1) = new_meta;
                    WORD(G_ID, 2) = ctr + 1u;     // flight-stream index, bits 0-19: a history does not take 2^20 Woodcock steps (mu_max > 0 is checked)
                    *reinterpret_cast<float4 *>(&GRP(G_POS)) = pos2[q];
                }
                c_steps += en2[q] ? 1u : 0u;
                st = moved ? ((st & clrq) | ((out ? P_REFILL : P_COLLIDE) << (4 * jj[q]))) : st;
            }

        } while (0);

        if (todo & 0x800000u) do {                                      // ======== COLLIDE
        MC_PHASE_SLOT(P_COLLIDE)
            if (!active) break;
            // ---------------- real collision, CBCT_real325im.cu:599-656 ----------------------
            uint4 id = GRP(G_ID);
            const float4 pos = *reinterpret_cast<float4 *>(&GRP(G_POS));
            uint32_t meta = id.y;
            const int nint = (meta >> 8) & 0xF, kE = meta & 0xFF, mat = (meta >> 12) & 0x7;
            if (nint >= sc.max_scatter) {                             // scatter budget used up
                if (RECORD) { P.fates[WORD(G_REC, 0)] = 5u | ((uint32_t)nint << 28); P.fate_e[WORD(G_REC, 0)] = pos.w; }
                st = (st & clr) | (P_REFILL << (4 * j));
                break;
            }
            if (P.collide_check) {
                const float2 cs = __ldg(sc.view_cs + (id.w >> 20));
                const float xr = pos.x * cs.x + pos.y * cs.y, yr = -pos.x * cs.y + pos.y * cs.x;
                if (RING ? (xr * xr + yr * yr >= sc.ring_r2 || fabsf(pos.z) >= sc.half)
                            : (xr >= sc.dod || fabsf(yr) >= sc.half || fabsf(pos.z) >= sc.half)) {        // :613-619
                    if (RECORD) { P.fates[WORD(G_REC, 0)] = 4u | ((uint32_t)nint << 28); P.fate_e[WORD(G_REC, 0)] = pos.w; }
                    st = (st & clr) | (P_REFILL << (4 * j));
                    break;
                }
            }
            meta += 0x100u;                                           // nint++
            c_int++;
            const uint2 re = philox2x32_10(id.x, (meta & 0xFF000000u) | (STREAM_EVENT << 22) | (id.z >> 20), P.key);
            id.y = meta; id.z += 0x100000u;
            GRP(G_ID) = id;
            const float u_sel = u01(re.x);
            const float4 tb = s_tab[mat * TAB_ROWS + kE];
            // photoelectric (:651-655) | coherent: no deflection (:656-695), or an angle from the form factor | Compton --
            // selects, not branches: the lanes of a visit take all three
            const bool is_pe = u_sel <= tb.y, is_coh = !is_pe & (u_sel <= tb.z), is_com = !is_pe & !is_coh;
            c_abs += is_pe ? 1u : 0u; c_coh += is_coh ? 1u : 0u; c_comp += is_com ? 1u : 0u;
            if (RECORD && is_pe) { P.fates[WORD(G_REC, 0)] = 3u | ((uint32_t)(nint + 1) << 28); P.fate_e[WORD(G_REC, 0)] = pos.w; }
            WORDF(G_DIR, 3) = u01(re.y);                              // the azimuthal variate of a deflection (read by COMPTON only)
            if (RAYLEIGH && is_coh) WORD(G_ID, 1) = meta | 0x10000u;
            const uint32_t next = is_pe ? P_REFILL : (is_com || (RAYLEIGH && is_coh)) ? P_COMPTON : P_STEP;
            st = (st & clr) | (next << (4 * j));

        } while (0);

        if (todo & 0x80000000u) do {                                    // ======== COMPTON
        MC_PHASE_SLOT(P_COMPTON)
            if (!active) break;
            // ---------------- one round of Kahn's method, :701-736 -----------------------------
            uint4 id = GRP(G_ID);
            const float E0 = WORDF(G_POS, 3);
            const float lam = __fdividef(511.0f, E0);
            const uint32_t ne = id.z >> 20;
            // one Philox block per round: r2, r3 from the high 23 bits of the two words, r1 (which only
            // picks the branch) from their 2 x 9 low bits -- disjoint bits, hence independent variates
            const uint2 ra = philox2x32_10(id.x, (id.y & 0xFF000000u) | (STREAM_EVENT << 22) | ne, P.key);
            id.z += 0x100000u;
            const float r2 = u01(ra.x), r3 = u01(ra.y);
            float cos_t, E;
            if (RAYLEIGH && (id.y & 0x10000u)) {
                // one round of form-factor sampling: x^2 from F^2 on [0, x^2_max], accepted with (1 + cos^2)/2
                const int rn = P.ray_n;
                const float *gx = s_ray + ((id.y >> 12) & 0x7u) * 2 * ray_stride, *gc = gx + ray_stride;
                const float xm = E0 * (1.0f / 12.3984193f), x2max = xm * xm;
                const float amax = ray_interp(gx, gc, rn, x2max);
                const float x2 = fminf(ray_interp(gc, gx, rn, r2 * amax), x2max);
                cos_t = fmaxf(1.0f - 2.0f * __fdividef(x2, x2max), -1.0f);
                if (!(r3 <= 0.5f * (1.0f + cos_t * cos_t))) { WORD(G_ID, 2) = id.z; break; }
                id.y &= ~0x10000u;
                E = E0;
            } else {
            const float r1 = ((float)(((ra.x & 0x1FFu) << 9) | (ra.y & 0x1FFu)) + 0.5f) * (1.0f / 262144.0f);
            const bool br1 = r1 * (9.0f * lam + 2.0f) < (lam + 2.0f);
            const float ro1 = 1.0f + __fdividef(2.0f, lam) * r2;
            const float ro2 = __fdividef(lam + 2.0f, lam + 2.0f * (1.0f - r2));
            const float ro = br1 ? ro1 : ro2;
            const float iro = __fdividef(1.0f, ro);
            const float t = lam - ro * lam + 1.0f;
            const float lim = br1 ? 4.0f * (iro - iro * iro) : 0.5f * (t * t + iro);
            if (!(r3 <= lim)) { WORD(G_ID, 2) = id.z; break; }     // rejected: next round next time
            const float lam_d = ro * lam;
            cos_t = 1.0f - (lam_d - lam);
            cos_t = fmaxf(cos_t, -1.0f);                              // :746-747
            E = __fdividef(511.0f, lam_d);
            }
            const float sin_t = sqrtf(fmaxf(0.f, 1.0f - cos_t * cos_t));
            const int kE = min(max((int)(E + 0.5f), 0), TAB_ROWS - 1);
            WORDF(G_POS, 3) = E;
            id.y = (id.y & ~0xFFu) | (uint32_t)kE;
            GRP(G_ID) = id;
            float4 dir = *reinterpret_cast<float4 *>(&GRP(G_DIR));
            float sphi, cphi;
            __sincosf(6.2831853071795865f * dir.w, &sphi, &cphi);     // phi = 2 pi u, :764 (MUFU, |err| ~1e-6)
            const float dx = dir.x, dy = dir.y, dz = dir.z;
            // direction update, :768-780, as a rotation of the unit vector.  With
            // (sin th_a cos ph_a, sin th_a sin ph_a, cos th_a) = d the reference's formulas are
            //   d' = cos_t d + sin_t (cos phi e1 + sin phi e2),
            //   e1 = (cos th_a cos ph_a, cos th_a sin ph_a, -sin th_a), e2 = (-sin ph_a, cos ph_a, 0).
            const float st2 = dx * dx + dy * dy;
            // (selects: a flight along +-z takes e1 = (dz, 0, 0), e2 = (0, 1, 0))
            const bool tilted = st2 > 1e-12f;
            const float ist = rsqrtf(tilted ? st2 : 1.0f), sta = st2 * ist;
            const float e1x = tilted ? dx * dz * ist : dz, e1y = tilted ? dy * dz * ist : 0.f, e1z = tilted ? -sta : 0.f;
            const float e2x = tilted ? -dy * ist : 0.f, e2y = tilted ? dx * ist : 1.f;
            const float a = sin_t * cphi, b = sin_t * sphi;
            const float nxd = cos_t * dx + a * e1x + b * e2x;
            const float nyd = cos_t * dy + a * e1y + b * e2y;
            const float nzd = cos_t * dz + a * e1z;
            const float nn = rsqrtf(nxd * nxd + nyd * nyd + nzd * nzd);
            dir.x = nxd * nn; dir.y = nyd * nn; dir.z = nzd * nn;
            *reinterpret_cast<float4 *>(&GRP(G_DIR)) = dir;
            if (CLEAR && P.coct) {                                    // DIRECTIONAL: the clearance depends on the new direction
                const float4 pc = *reinterpret_cast<float4 *>(&GRP(G_POS));
                int ix = __float_as_int(fmaf(pc.x, sc.inv_pitch, vox_off[0]) + 12582912.0f) - 0x4B400000;
                int iy = __float_as_int(fmaf(pc.y, sc.inv_pitch, vox_off[1]) + 12582912.0f) - 0x4B400000;
                int iz = __float_as_int(fmaf(pc.z, sc.inv_pitch, vox_off[2]) + 12582912.0f) - 0x4B400000;
                ix = min(max(ix, 0), sc.nx - 1); iy = min(max(iy, 0), sc.ny - 1); iz = min(max(iz, 0), sc.nz - 1);
                const unsigned oc = (dir.x > 0.f ? 1u : 0u) | (dir.y > 0.f ? 2u : 0u) | (dir.z > 0.f ? 4u : 0u);
                const uint32_t qd = __ldg(P.clear + oc * P.coct + ((unsigned)((iz >> P.cshift) * P.cgy + (iy >> P.cshift)) * (unsigned)P.cgx + (unsigned)(ix >> P.cshift)));
                WORD(G_ID, 1) = (id.y & ~0xFE0000u) | (qd << 17);
            }
            st = (st & clr) | (P_STEP << (4 * j));

        } while (0);

        if (todo & 0x80u) do {                                          // ======== REFILL
        MC_PHASE_SLOT(P_REFILL)
        // ---------------- phase == P_REFILL: finish the ended history, start the next one ----------
        // unit bookkeeping is warp-uniform: executed by every lane
        if (next_off >= unit_cnt && !grid_done) {
            unsigned long long unit = 0;
            if (lane == 0) unit = atomicAdd(P.work, 1ull);
            unit = __shfl_sync(0xffffffffu, unit, 0);
            if (unit >= P.n_units) grid_done = true;
            else {
                const unsigned long long base = unit * (unsigned long long)MC_UNIT;
                unit_cnt = (uint32_t)min((unsigned long long)MC_UNIT, P.total - base);
                const uint32_t pv0 = (uint32_t)(base / P.cnt);
                n0 = (uint32_t)(base - (unsigned long long)pv0 * P.cnt);
                v0 = pv0 / npix; p0 = pv0 - v0 * npix;
                pi0 = p0 / (uint32_t)sc.det_nx; pj0 = p0 - pi0 * (uint32_t)sc.det_nx;
                next_off = 0;
            }
        }
        const unsigned m_ref = __ballot_sync(0xffffffffu, active);
        // the cached source ray of this warp: read by all lanes together, right after the ballot and before anything
        // divergent, so that the one lane that may refresh it further down cannot overtake a reader
        __syncwarp();
        const uint4 sca = s_src[0], scb = s_src[1], scc = s_src[2];
        const uint32_t first_off = next_off;
        const uint32_t my_off = next_off + __popc(m_ref & lt_mask);
        next_off = min(next_off + (uint32_t)__popc(m_ref), unit_cnt);
        if (!active) break;
        {
            const uint4 id = GRP(G_ID);
            if (id.y & 0x8000u) {                        // the slot's history left the volume on its last step
                WORD(G_ID, 1) = id.y & ~0x8000u;
                const int nint = (id.y >> 8) & 0xF;
                const float4 pos = *reinterpret_cast<float4 *>(&GRP(G_POS));
                if (nint == 0) {                         // unscattered: lands in the pixel it was aimed at (:567-590)
                    const uint32_t pvw = id.w;
                    const uint32_t pvq = (pvw >> 20) * npix + (pvw & 0xFFFFFu);
                    if (pvq != cur_pv) {
                        if (prim_cnt) { atomicAdd(P.image0 + cur_pv, (int)prim_cnt); atomicAdd(P.image5 + cur_pv, (int)prim_cnt); }
                        cur_pv = pvq; prim_cnt = 0;
                    }
                    prim_cnt += sc.eid ? (uint32_t)(pos.w * (float)MONTE_MC_EID_SCALE + 0.5f) : 1u; c_prim++;
                    e_prim += (unsigned long long)(pos.w * 1024.f + 0.5f);
                    if (RECORD) { P.fates[WORD(G_REC, 0)] = 1u | ((pvw & 0xFFFFFu) << 8); P.fate_e[WORD(G_REC, 0)] = pos.w; }
                } else {                                 // scatter detection, CBCT_real325im.cu:823-843
                uint32_t fate = 4u | ((uint32_t)nint << 28);
                const float4 dir = *reinterpret_cast<float4 *>(&GRP(G_DIR));
                const int view = (int)(id.w >> 20);
                const float2 cs = __ldg(sc.view_cs + view);
                const float xr = pos.x * cs.x + pos.y * cs.y, yr = -pos.x * cs.y + pos.y * cs.x;      // rotate by -beta
                const float dxr = dir.x * cs.x + dir.y * cs.y, dyr = -dir.x * cs.y + dir.y * cs.x;
                if (RING) {
                    // the flight leaves through the cylinder x^2 + y^2 = R^2 (circle3_2.cpp:243-252, a ring of angular bins)
                    const float qa = dxr * dxr + dyr * dyr, qb = xr * dxr + yr * dyr, qc = xr * xr + yr * yr - sc.ring_r2;
                    const float disc = qb * qb - qa * qc;
                    const float xf = fmaf(1000.f, dxr, xr), yf = fmaf(1000.f, dyr, yr);
                    if (qa > 0.f && disc >= 0.f && xf * xf + yf * yf >= sc.ring_r2) {
                        // the root where the line LEAVES the cylinder; negative when the last Woodcock step ended beyond the
                        // ring (the history came from inside: that crossing is the hit either way)
                        const float t = __fdividef(sqrtf(disc) - qb, qa);
                        const float zd = fmaf(t, dir.z, pos.z);
                        if (fabsf(zd) <= sc.half) {
                            float phi = atan2f(fmaf(t, dyr, yr), fmaf(t, dxr, xr));
                            if (phi < 0.f) phi += 6.2831853071795865f;
                            const int by = min((int)(phi * sc.ring_kphi), sc.det_ny - 1), bx = (int)((sc.half - zd) * sc.inv_pixel);
                            if (by >= 0 && bx >= 0 && bx < sc.det_nx) {
                                const uint32_t bin = (uint32_t)(by * sc.det_nx + bx);
                                atomicAdd(P.image5 + (size_t)view * npix + bin, sc.eid ? (int)(pos.w * (float)MONTE_MC_EID_SCALE + 0.5f) : 1);
                                c_scat++;
                                e_scat += (unsigned long long)(pos.w * 1024.f + 0.5f);
                                fate = 2u | (bin << 8) | ((uint32_t)nint << 28);
                            }
                        }
                    }
                } else
                if (dxr > 0.f) {
                    const float t = __fdividef(sc.dod - xr, dxr);
                    const float yd = fmaf(t, dyr, yr), zd = fmaf(t, dir.z, pos.z);
                    if (fabsf(yd) <= sc.half && fabsf(zd) <= sc.half && fmaf(1000.f, dxr, xr) >= sc.dod) {
                        const int by = (int)((sc.half - yd) * sc.inv_pixel), bx = (int)((sc.half - zd) * sc.inv_pixel);
                        if (by >= 0 && by < sc.det_ny && bx >= 0 && bx < sc.det_nx) {
                            const uint32_t bin = (uint32_t)(by * sc.det_nx + bx);
                            atomicAdd(P.image5 + (size_t)view * npix + bin, sc.eid ? (int)(pos.w * (float)MONTE_MC_EID_SCALE + 0.5f) : 1);
                            c_scat++;
                            e_scat += (unsigned long long)(pos.w * 1024.f + 0.5f);
                            fate = 2u | (bin << 8) | ((uint32_t)nint << 28);
                        }
                    }
                }
                if (RECORD) { P.fates[WORD(G_REC, 0)] = fate; P.fate_e[WORD(G_REC, 0)] = pos.w; }
                }
            }
        }
        if (my_off >= unit_cnt) {                        // no history left in this unit for me
            if (grid_done) st &= clr;                    // ... nor anywhere: the slot is done
            break;                                    // else: stays REFILL, the next visit opens a new unit
        }
        {
            // history index -> (pixel, photon): divisions only in the rare general case
            uint32_t n = n0 + my_off, dpv;
            if (few_pixels_per_unit) { dpv = 0u; while (n >= P.cnt) { n -= P.cnt; dpv++; } }
            else { dpv = n / P.cnt; n -= dpv * P.cnt; }
            uint32_t pix = p0 + dpv, vrel = v0, pi = pi0, pj = pj0 + dpv;
            if (pix >= npix || pj >= (uint32_t)sc.det_nx) {
                if (pix >= npix) { const uint32_t q = pix / npix; vrel += q; pix -= q * npix; }
                pi = pix / (uint32_t)sc.det_nx; pj = pix - pi * (uint32_t)sc.det_nx;
            }
            const int view = P.view_begin + (int)vrel;
            const uint32_t pva = (uint32_t)view * npix + pix;
            n += P.n_begin;
            const unsigned long long hid = (unsigned long long)pva * P.per + n;
            const uint32_t c0 = (uint32_t)hid;
            const uint32_t c1hi = ((uint32_t)(hid >> 32) & 0xFFu) << 24;
            uint32_t rec_idx = 0;
            if (RECORD) { rec_idx = pix * P.per + n; WORD(G_REC, 0) = rec_idx; }
            c_hist++;
            // ---- source, CBCT_real325im.cu:464-540 (exact aim at the pixel) ----
            float E, dx, dy, dz, ex, ey, ez;
            int kE;
            uint32_t qe = 0u;
            bool miss;
            if (src_cached && scc.x == pva) {                 // same pixel as the cached ray: nothing to compute
                ex = __uint_as_float(sca.x); ey = __uint_as_float(sca.y); ez = __uint_as_float(sca.z); E = __uint_as_float(sca.w);
                dx = __uint_as_float(scb.x); dy = __uint_as_float(scb.y); dz = __uint_as_float(scb.z);
                miss = scb.w != 0u; qe = scc.y; kE = (int)scc.z;
            } else {
            float uy = 0.5f, uz = 0.5f;
            if (sc.source_mode == MONTE_MC_SOURCE_CONE) {
                const uint2 r = philox2x32_10(c0, c1hi | (STREAM_SOURCE << 22), P.key);
                uy = u01(r.x); uz = u01(r.y);
            }
            E = sc.mono_keV;
            if (sc.n_bins > 0) {                              // CBCT_real325im.cu:492-498
                const uint2 r = philox2x32_10(c0, c1hi | (STREAM_SOURCE << 22) | 1u, P.key);
                const float ue = u01(r.x);
                int lo = 0, hi = sc.n_bins;                  // first k with ue <= cdf[k+1]
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (ue <= s_cdf[mid + 1]) hi = mid; else lo = mid + 1; }
                if (lo < sc.n_bins) E = (float)(lo + 1) * sc.bin_keV;
            }
            kE = min(max((int)(E + 0.5f), 0), TAB_ROWS - 1);
            const float zl = sc.half - sc.pixel * ((float)pj + uz);
            const float2 cs = __ldg(sc.view_cs + view);
            float dxr, dyr, rn, sx, sy;
            if (RING) {                                       // aim from the origin at bin (pi, pj) of the ring
                float sph, cph;
                sincosf(sc.ring_dphi * ((float)pi + uy), &sph, &cph);
                rn = rsqrtf(sc.ring_r2 + zl * zl);
                dxr = sc.ring_r * cph * rn; dyr = sc.ring_r * sph * rn;
                sx = 0.f; sy = 0.f;
            } else {
                const float yl = sc.half - sc.pixel * ((float)pi + uy);
                rn = rsqrtf(sc.dsd * sc.dsd + yl * yl + zl * zl);
                dxr = sc.dsd * rn; dyr = yl * rn;
                sx = -sc.dso * cs.x; sy = -sc.dso * cs.y;
            }
            dx = dxr * cs.x - dyr * cs.y;
            dy = dxr * cs.y + dyr * cs.x;
            dz = zl * rn;
            // analytic flight to the clip box: branch-free slab method (a zero direction component gives
            // +-inf bounds, which fminf/fmaxf handle; CUDA's fminf/fmaxf drop a NaN operand)
            float t0 = 0.f, t1 = 1e30f;
            {
                const float o3[3] = {sx, sy, 0.f}, d3[3] = {dx, dy, dz};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float inv = __fdividef(1.0f, d3[a]);
                    const float ta = (sc.clip_lo[a] - o3[a]) * inv, tb = (sc.clip_hi[a] - o3[a]) * inv;
                    t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
                }
            }
            miss = t0 >= t1;
            ex = fmaf(t0, dx, sx); ey = fmaf(t0, dy, sy); ez = t0 * dz;
            if (CLEAR && !miss) {                             // clearance of the cell the photon enters through
                // the entry point lies exactly on a voxel face of the clip box: take the voxel 1e-3 of a voxel
                // side further along the ray, so that the choice does not hang on rounding (the oracle does the same)
                int ix = __float_as_int(fmaf(ex, sc.inv_pitch, vox_off[0]) + 1e-3f * dx + 12582912.0f) - 0x4B400000;
                int iy = __float_as_int(fmaf(ey, sc.inv_pitch, vox_off[1]) + 1e-3f * dy + 12582912.0f) - 0x4B400000;
                int iz = __float_as_int(fmaf(ez, sc.inv_pitch, vox_off[2]) + 1e-3f * dz + 12582912.0f) - 0x4B400000;
                ix = min(max(ix, 0), sc.nx - 1); iy = min(max(iy, 0), sc.ny - 1); iz = min(max(iz, 0), sc.nz - 1);
                const unsigned oc = (dx > 0.f ? 1u : 0u) | (dy > 0.f ? 2u : 0u) | (dz > 0.f ? 4u : 0u);
                qe = __ldg(P.clear + oc * P.coct + ((unsigned)((iz >> P.cshift) * P.cgy + (iy >> P.cshift)) * (unsigned)P.cgx + (unsigned)(ix >> P.cshift)));
            }
            // the first lane of the visit refreshes the cache (one writer: entries are never mixed; readers took their copy
            // right after the ballot)
            if (src_cached && my_off == first_off) {
                s_src[0] = make_uint4(__float_as_uint(ex), __float_as_uint(ey), __float_as_uint(ez), __float_as_uint(E));
                s_src[1] = make_uint4(__float_as_uint(dx), __float_as_uint(dy), __float_as_uint(dz), miss ? 1u : 0u);
                s_src[2] = make_uint4(pva, qe, (uint32_t)kE, 0u);
            }
            }
            if (miss) {                                       // misses the phantom: unscattered
                if (pva != cur_pv) {
                    if (prim_cnt) { atomicAdd(P.image0 + cur_pv, (int)prim_cnt); atomicAdd(P.image5 + cur_pv, (int)prim_cnt); }
                    cur_pv = pva; prim_cnt = 0;
                }
                prim_cnt += sc.eid ? (uint32_t)(E * (float)MONTE_MC_EID_SCALE + 0.5f) : 1u; c_prim++;
                e_prim += (unsigned long long)(E * 1024.f + 0.5f);
                if (RECORD) { P.fates[rec_idx] = 1u | (pix << 8); P.fate_e[rec_idx] = E; }
                // the slot stays in REFILL and takes another history on the next visit
            } else {
                *reinterpret_cast<float4 *>(&GRP(G_POS)) = make_float4(ex, ey, ez, E);
                *reinterpret_cast<float4 *>(&GRP(G_DIR)) = make_float4(dx, dy, dz, 0.f);
                GRP(G_ID) = make_uint4(c0, c1hi | (qe << 17) | (uint32_t)kE, 0u, ((uint32_t)view << 20) | pix);
                st = (st & clr) | (P_STEP << (4 * j));
            }
        }
        } while (0);
#undef MC_PHASE_SLOT
    }
#undef GRP
#undef WORD
#undef WORDF
    if (prim_cnt) { atomicAdd(P.image0 + cur_pv, (int)prim_cnt); atomicAdd(P.image5 + cur_pv, (int)prim_cnt); }
    if (P.stats) {
        const uint32_t v[8] = {c_hist, c_prim, c_scat, c_abs, c_int, c_coh, c_comp, c_steps};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            // 32-bit per-lane counters, summed in 64 bits
            unsigned long long s = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && s) atomicAdd(P.stats + i, s);
        }
        unsigned long long ep = e_prim, es = e_scat;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ep += __shfl_xor_sync(0xffffffffu, ep, o); es += __shfl_xor_sync(0xffffffffu, es, o); }
        if (lane == 0) { if (ep) atomicAdd(P.stats + ST_EPRIM, ep); if (es) atomicAdd(P.stats + ST_ESCAT, es); }
    }
}

// counts -> -ln(I/I0), CBCT_real325im.cu:267-285
// (-log(int) is the double overload there, log(float(per)) the float one: log_per is computed by
// the host's logf so both terms round exactly as in the reference)
__global__ void counts_to_map_kernel(const int32_t *counts, size_t n, int32_t per, float log_per, float *map) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = counts[i];
    c = min(c, per);
    if (c == 0) c = 1;
    map[i] = (float)(-log((double)c) + (double)log_per);
}

// Multi-device tally reduction fused with the counts -> map epilogue (SURVEY 8e + row A12): ONE kernel on the root
// device sums the int32 tallies of up to MAX_DEV - 1 peer devices straight out of their memory (NVLink P2P loads,
// 128-bit, coalesced) into its own, and -- when maps are wanted -- applies CBCT_real325im.cu:267-285 to the sum in
// the same pass.  n covers [image0 | image5] of the requested views, so there is one launch per call.
struct TallyPeers {
    const int32_t *p[MAX_DEV];
    int n;
};
__global__ void tally_reduce_map_kernel(const TallyPeers peers, int32_t *counts, size_t n, int32_t per, float log_per, float *map) {
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    int c[4];
    if (i4 + 4 <= n) {
        const int4 v = *reinterpret_cast<const int4 *>(counts + i4);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
        for (int k = 0; k < peers.n; k++) {
            const int4 w = *reinterpret_cast<const int4 *>(peers.p[k] + i4);
            c[0] += w.x; c[1] += w.y; c[2] += w.z; c[3] += w.w;
        }
        if (peers.n) *reinterpret_cast<int4 *>(counts + i4) = make_int4(c[0], c[1], c[2], c[3]);
    } else {
        for (int j = 0; j < 4; j++) {
            if (i4 + j >= n) { c[j] = 1; continue; }
            c[j] = counts[i4 + j];
            for (int k = 0; k < peers.n; k++) c[j] += peers.p[k][i4 + j];
            if (peers.n) counts[i4 + j] = c[j];
        }
    }
    if (!map) return;
    float m[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int q = min(c[j], per);
        if (q == 0) q = 1;
        m[j] = (float)(-log((double)q) + (double)log_per);
    }
    if (i4 + 4 <= n) *reinterpret_cast<float4 *>(map + i4) = make_float4(m[0], m[1], m[2], m[3]);
    else for (int j = 0; j < 4 && i4 + j < n; j++) map[i4 + j] = m[j];
}

// Label volume of a multi-device call: every device uploads 1/N of the labels over its own PCIe link, then completes
// its copy by loading the other parts out of the peers' memory over NVLink (N x 34 MB through the host's memory system
// took 1.5 ms of an 8 ms C2 step on 8 B200s; 34 MB once + NVLink is ~0.25 ms).  Parts are cut on 16-byte words.
struct LabelPeers {
    const uint4 *p[8];          // the peers' label buffers (index = device)
    unsigned long long w_end[8];   // part j = words [w_end[j-1], w_end[j])
    int n, self;
};
__global__ void label_gather_kernel(const LabelPeers L, uint4 *mine, unsigned long long n_words) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        int j = 0;
        while (j < L.n - 1 && w >= L.w_end[j]) j++;
        if (j != L.self) mine[w] = L.p[j][w];
    }
}

}  // namespace monte

using namespace monte;

// ------------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------------
constexpr unsigned MC_WORK_RING = 16;
struct monte_mc_scene {
    McSceneDev dev;
    monte_mc_geom geom;
    void *d_labels = nullptr;
    void *d_small = nullptr, *h_small = nullptr;                       // all small tables: one device block + its pinned staging copy
    void *d_tab = nullptr, *d_inv = nullptr, *d_cdf = nullptr, *d_view = nullptr, *d_ray = nullptr;   // (inside d_small)
    size_t cap_labels = 0, cap_small = 0;                              // grow-only capacities (bytes)
    int ray_n = 0;                                                     // > 0: form-factor tables uploaded (coherent_mode 1)
    // two-level majorant (tracking_mode CLEARANCE): clearance grid, light-majorant table, the material it excludes
    void *d_clear = nullptr, *d_invlo = nullptr;                        // (d_invlo inside d_small)
    size_t cap_clear = 0;
    int heavy = -1, cshift = 0, cg[3] = {0, 0, 0};
    unsigned coct = 0;                                                 // DIRECTIONAL: cells per octant grid
    float cunit = 0.f;
    monte_mc_volume vol;                                               // kept for scene_update_labels
    int n_mat_host = 0;
    uint64_t clear_hash = 0;                                           // labels the resident grid was built from
    int clear_key[6] = {0, 0, 0, -1, -1, -1};                          // nx, ny, nz, cell_log2, heavy, n_materials
    uint64_t labels_hash = 0;                                          // content hash of the resident labels (0: unknown)
    size_t labels_n = 0;
    // majorant_mode PRESENT: materials that occur in the resident labels (bit m = material m); what scene_update_labels
    // needs to rebuild the tables when that set changes
    uint32_t present_mask = 0xffffffffu;
    bool present_valid = false;
    monte_mc_xs *xs_host = nullptr;
    bool adaptive_host = false;
    size_t o_tab = 0, o_inv = 0, o_invlo = 0, tables_bytes = 0;          // table region inside d_small / h_small
    // work-unit counters: one per launch out of a ring, so that launches of one scene that overlap on different
    // streams do not share (and re-zero) a counter.  More than MC_WORK_RING launches of one scene in flight at
    // once are not supported.
    unsigned long long *d_work = nullptr;
    mutable unsigned work_next = 0;
    size_t smem = 0;
    size_t h2d_bytes = 0;
};

static int check_mc(const monte_mc_geom *g, const monte_mc_volume *vol, const monte_mc_xs *xs) {
    MONTE_ARG(g && vol && xs, "mc: NULL argument");
    MONTE_ARG(g->n_views > 0 && g->ny > 0 && g->nx > 0 && g->pixel > 0, "mc: bad detector geometry");
    MONTE_ARG((uint64_t)g->n_views * g->ny * g->nx < (1ull << 32), "mc: views*pixels must fit 32 bits");
    MONTE_ARG((size_t)g->ny * g->nx < (1u << 20), "mc: detector has more than 2^20 pixels");
    MONTE_ARG(g->max_scatter >= 0 && g->max_scatter <= 15, "mc: max_scatter must be 0..15");
    MONTE_ARG(g->detector_mode == MONTE_MC_DETECTOR_COUNTING || g->detector_mode == MONTE_MC_DETECTOR_ENERGY,
              "mc: unknown detector_mode %d", g->detector_mode);
    MONTE_ARG(g->detector_shape == MONTE_MC_DETECTOR_FLAT || g->detector_shape == MONTE_MC_DETECTOR_RING,
              "mc: unknown detector_shape %d", g->detector_shape);
    if (g->detector_shape == MONTE_MC_DETECTOR_RING) {
        MONTE_ARG(g->ring_radius > 0, "mc: detector_shape RING needs ring_radius > 0 (got %g)", g->ring_radius);
        const double rx = fmax(fabs(vol->clip_lo[0]), fabs(vol->clip_hi[0])), ry = fmax(fabs(vol->clip_lo[1]), fabs(vol->clip_hi[1]));
        MONTE_ARG(sqrt(rx * rx + ry * ry) < g->ring_radius, "mc: the clip box must lie inside the detector ring (corner at %g cm, ring_radius %g)",
                  sqrt(rx * rx + ry * ry), g->ring_radius);
    }
    MONTE_ARG(vol->nx > 0 && vol->ny > 0 && vol->nz > 0 && vol->pitch > 0, "mc: bad volume");
    MONTE_ARG((uint64_t)vol->nx * vol->ny * vol->nz < (1ull << 32), "mc: label volume has 2^32 voxels or more");
    MONTE_ARG(vol->tracking_mode >= MONTE_MC_TRACK_GLOBAL && vol->tracking_mode <= MONTE_MC_TRACK_DIRECTIONAL,
              "mc: unknown tracking_mode %d", vol->tracking_mode);
    MONTE_ARG((vol->tracking_mode != MONTE_MC_TRACK_CLEARANCE && vol->tracking_mode != MONTE_MC_TRACK_ADAPTIVE &&
               vol->tracking_mode != MONTE_MC_TRACK_DIRECTIONAL) ||
              (vol->clearance_cell_log2 >= 0 && vol->clearance_cell_log2 <= 8),
              "mc: clearance_cell_log2 must be 0..8 (got %d)", vol->clearance_cell_log2);
    MONTE_ARG(vol->majorant_mode == MONTE_MC_MAJORANT_ALL || vol->majorant_mode == MONTE_MC_MAJORANT_PRESENT,
              "mc: unknown majorant_mode %d", vol->majorant_mode);
    MONTE_ARG(xs->n_materials >= 1 && xs->n_materials <= MONTE_MC_MAX_MATERIALS, "mc: n_materials must be 1..%d", MONTE_MC_MAX_MATERIALS);
    MONTE_ARG(g->coherent_mode == MONTE_MC_COHERENT_FORWARD || g->coherent_mode == MONTE_MC_COHERENT_FORMFACTOR,
              "mc: unknown coherent_mode %d", g->coherent_mode);
    if (g->coherent_mode == MONTE_MC_COHERENT_FORMFACTOR) {
        MONTE_ARG(xs->ff_points >= 2 && xs->ff_points <= MONTE_MC_FF_POINTS,
                  "mc: coherent_mode FORMFACTOR needs form-factor tables (ff_points = %d, must be 2..%d)", xs->ff_points, MONTE_MC_FF_POINTS);
        for (int m = 0; m < xs->n_materials; m++) {
            MONTE_ARG(xs->ff_x2[m][0] == 0.f && xs->ff_cum[m][0] == 0.f, "mc: form-factor table of material %d must start at (0, 0)", m);
            for (int i = 1; i < xs->ff_points; i++)
                MONTE_ARG(xs->ff_x2[m][i] > xs->ff_x2[m][i - 1] && xs->ff_cum[m][i] >= xs->ff_cum[m][i - 1],
                          "mc: form-factor table of material %d is not ascending at point %d", m, i);
            MONTE_ARG(xs->ff_cum[m][xs->ff_points - 1] > 0.f, "mc: form-factor table of material %d is all zero", m);
        }
    }
    const int dims[3] = {vol->nx, vol->ny, vol->nz};
    for (int a = 0; a < 3; a++) {
        MONTE_ARG(vol->clip_lo[a] < vol->clip_hi[a], "mc: empty clip box");
        const double tol = 1e-6 * vol->pitch;
        MONTE_ARG(vol->clip_lo[a] >= vol->origin[a] - tol && vol->clip_hi[a] <= vol->origin[a] + dims[a] * vol->pitch + tol,
                  "mc: clip box must lie inside the label volume (axis %d)", a);
    }
    return MONTE_OK;
}

// energies must index the tables (1..200 keV) and the majorant must be positive wherever a photon can be: a zero
// majorant makes every Woodcock step zero-length and the persistent kernel would never finish
static int check_spectrum(const monte_mc_xs *xs, const monte_mc_spectrum *spec) {
    double e_max = 140.0;                                              // spec == NULL: the shipped 140 keV
    if (spec) {
        MONTE_ARG(spec->n_bins >= 0 && spec->n_bins <= 4096, "mc: bad spectrum (n_bins = %d)", spec->n_bins);
        if (spec->n_bins == 0) {
            MONTE_ARG(spec->mono_keV > 0 && spec->mono_keV <= MONTE_MC_TABLE_ROWS - 1, "mc: mono_keV must be in (0, %d] (got %g)",
                      MONTE_MC_TABLE_ROWS - 1, spec->mono_keV);
            e_max = spec->mono_keV;
        } else {
            MONTE_ARG(spec->cdf != nullptr, "mc: spectrum has bins but no cdf");
            MONTE_ARG(spec->bin_keV > 0 && spec->n_bins * spec->bin_keV <= MONTE_MC_TABLE_ROWS - 1,
                      "mc: spectrum must end at or below %d keV (n_bins * bin_keV = %g)", MONTE_MC_TABLE_ROWS - 1, spec->n_bins * spec->bin_keV);
            e_max = spec->n_bins * spec->bin_keV;
            if (spec->mono_keV > e_max) e_max = spec->mono_keV;        // fallback energy of an incomplete cdf (:497)
            MONTE_ARG(e_max <= MONTE_MC_TABLE_ROWS - 1, "mc: mono_keV (cdf fallback) above %d keV", MONTE_MC_TABLE_ROWS - 1);
        }
    }
    // Compton scattering only lowers the energy: rows 1 .. round(e_max) are reachable (row 0 only below 0.5 keV,
    // where the tables end; it is checked too because the kernel clamps to it)
    const int k_hi = (int)(e_max + 0.5);
    for (int k = 0; k <= k_hi && k < MONTE_MC_TABLE_ROWS; k++) {
        double mumax = 0;
        for (int m = 0; m < xs->n_materials; m++) mumax = fmax(mumax, (double)xs->total[m][k] * (double)xs->density[m]);
        MONTE_ARG(mumax > 0 && mumax < 1e30, "mc: the Woodcock majorant is %g at %d keV (tables must be positive up to the highest source energy)", mumax, k);
    }
    return MONTE_OK;
}

extern "C" {

static int grow(void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return MONTE_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    MONTE_CUDA(cudaMalloc(p, bytes));
    *cap = bytes;
    return MONTE_OK;
}

// 64-bit content hash of the label volume (four interleaved multiply-xorshift lanes, ~2 ms for 325^3): the
// host-buffer entry point re-uploads the scene on every call, and rebuilding the clearance grid (25-110 ms) for
// labels that have not changed would cost more than the transport itself
static uint64_t hash_labels_1(const uint8_t *p, size_t n) {
    uint64_t h[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t w[4];
        memcpy(w, p + i, 32);
        for (int l = 0; l < 4; l++) { h[l] = (h[l] ^ w[l]) * 0x9FB21C651E98DF25ull; h[l] ^= h[l] >> 32; }
    }
    for (; i < n; i += 8) {                            // at most 31 bytes remain: folded in 8-byte pieces
        uint64_t w = 0;
        memcpy(&w, p + i, n - i < 8 ? n - i : 8);
        h[0] = (h[0] ^ w) * 0x9FB21C651E98DF25ull; h[0] ^= h[0] >> 32;
    }
    uint64_t r = n;
    for (int l = 0; l < 4; l++) { r = (r ^ h[l]) * 0xD6E8FEB86659FD93ull; r ^= r >> 29; }
    return r;
}

// the same on eight host threads for volumes of 8 MB and more (325^3: 2 ms -> ~0.4 ms)
static uint64_t hash_labels(const uint8_t *p, size_t n) {
    constexpr int T = 8;
    if (n < (8u << 20)) return hash_labels_1(p, n);
    uint64_t h[T];
    std::thread th[T];
    const size_t part = (n / T) & ~(size_t)63;
    for (int t = 0; t < T; t++) {
        const uint8_t *q = p + part * t;
        const size_t m = t == T - 1 ? n - part * t : part;
        th[t] = std::thread([q, m, &h, t] { h[t] = hash_labels_1(q, m); });
    }
    uint64_t r = n;
    for (int t = 0; t < T; t++) { th[t].join(); r = (r ^ h[t]) * 0xD6E8FEB86659FD93ull; r ^= r >> 29; }
    return r ? r : 1;
}

// majorant_mode PRESENT: which materials occur in the labels (bit m = material m = label m + 1; labels above n_mat use
// the last material, as in the transport).  One pass over the volume on eight host threads (325^3: ~2 ms), done when
// labels are uploaded, not per launch.
static uint32_t label_presence_1(const uint8_t *p, size_t n, int n_mat) {
    uint32_t seen[4] = {0, 0, 0, 0};                   // four independent accumulators (the shifts do not serialise)
    size_t i = 0;
    for (; i + 4 <= n; i += 4)
        for (int l = 0; l < 4; l++) seen[l] |= 1u << (p[i + l] < 31 ? p[i + l] : 31);
    for (; i < n; i++) seen[0] |= 1u << (p[i] < 31 ? p[i] : 31);
    const uint32_t raw = seen[0] | seen[1] | seen[2] | seen[3];       // bit L = label L occurs (31 = any label >= 31)
    uint32_t mask = 0;
    for (int L = 1; L < 32; L++) if (raw >> L & 1u) mask |= 1u << ((L <= n_mat ? L : n_mat) - 1);
    return mask;
}
static uint32_t label_presence(const uint8_t *p, size_t n, int n_mat) {
    constexpr int T = 8;
    if (n < (4u << 20)) return label_presence_1(p, n, n_mat);
    uint32_t m[T];
    std::thread th[T];
    const size_t part = n / T;
    for (int t = 0; t < T; t++) {
        const uint8_t *q = p + part * t;
        const size_t cnt = t == T - 1 ? n - part * t : part;
        th[t] = std::thread([q, cnt, &m, t, n_mat] { m[t] = label_presence_1(q, cnt, n_mat); });
    }
    uint32_t r = 0;
    for (int t = 0; t < T; t++) { th[t].join(); r |= m[t]; }
    return r;
}

// per-keV tables of a scene into its pinned staging block: Woodcock majorant over the materials in `mask`
// (CBCT_real325im.cu:867-868 takes every table it loaded: mask = all) and the branching ratios (:651,656)
static void fill_tables(monte_mc_scene *s, const monte_mc_xs *xs, uint32_t mask, bool adaptive) {
    const int nm = xs->n_materials;
    char *hs = (char *)s->h_small;
    float4 *tab = (float4 *)(hs + s->o_tab);
    float *inv = (float *)(hs + s->o_inv), *invlo = (float *)(hs + s->o_invlo);   // invlo: 1/mu_light, then the clearance thresholds
    for (int k = 0; k < TAB_ROWS; k++) {
        double mumax = 0, mulo = 0;
        for (int m = 0; m < nm; m++) {
            if (!(mask >> m & 1u)) continue;
            mumax = fmax(mumax, (double)xs->total[m][k] * (double)xs->density[m]);
            if (m != s->heavy) mulo = fmax(mulo, (double)xs->total[m][k] * (double)xs->density[m]);
        }
        // nothing that attenuates at this energy (an all-air volume under MAJORANT_PRESENT): the medium is transparent,
        // one step of 1e30 cm leaves the clip box (a zero step length would keep the persistent kernel spinning)
        inv[k] = mumax > 0 ? (float)(1.0 / mumax) : 1e30f;
        invlo[k] = mulo > 0 ? (float)(1.0 / mulo) : inv[k];
        // ADAPTIVE: light majorant only where a cut at D is less likely than a virtual collision, exp(-mu_light D) < 1 - mu_light/mu_max
        invlo[TAB_ROWS + k] = !adaptive ? 0.f : (mulo > 0 && mulo < mumax ? (float)(-log(1.0 - mulo / mumax) / mulo) : 1e30f);
        for (int m = 0; m < nm; m++) {
            const double mu = (double)xs->total[m][k];
            float4 t;
            t.x = mumax > 0 ? (float)fmin(1.0, (mu * (double)xs->density[m]) / mumax) : 0.f;   // (an absent material is never looked up)
            t.y = mu > 0 ? (float)((double)xs->photo[m][k] / mu) : 1.f;
            t.z = mu > 0 ? (float)(((double)xs->photo[m][k] + (double)xs->coh[m][k]) / mu) : 1.f;
            t.w = mulo > 0 ? (float)fmin(1.0, (mu * (double)xs->density[m]) / mulo) : t.x;   // acceptance against the light majorant
            tab[(size_t)m * TAB_ROWS + k] = t;
        }
    }
}

// clearance grid of the current labels (host transform) -> device; skipped when the resident grid was built from
// the same labels, geometry and tables
static int upload_clearance(monte_mc_scene *s, const uint8_t *labels, cudaStream_t st, uint64_t known_hash = 0) {
    const monte_mc_volume &v = s->vol;
    const uint64_t hsh = known_hash ? known_hash : hash_labels(labels, (size_t)v.nx * v.ny * v.nz);
    const bool oct = v.tracking_mode == MONTE_MC_TRACK_DIRECTIONAL;
    const int key[6] = {v.nx, v.ny, v.nz, v.clearance_cell_log2 + (oct ? 100 : 0), s->heavy, s->n_mat_host};
    int32_t d[3];
    if (int rc = monte_mc_clearance_dims(&v, v.clearance_cell_log2, d)) return rc;
    s->cshift = v.clearance_cell_log2;
    s->cunit = (float)(0.5 * (double)(1 << v.clearance_cell_log2) * v.pitch);
    if (s->d_clear && hsh == s->clear_hash && memcmp(key, s->clear_key, sizeof(key)) == 0) return MONTE_OK;
    const size_t ncell = (size_t)d[0] * d[1] * d[2];
    s->coct = oct ? (unsigned)ncell : 0u;
    std::vector<uint8_t> grid(ncell * (oct ? 8 : 1));
    if (int rc = oct ? monte_mc_clearance_grid_octants(&v, labels, s->n_mat_host, s->heavy, v.clearance_cell_log2, grid.data())
                     : monte_mc_clearance_grid(&v, labels, s->n_mat_host, s->heavy, v.clearance_cell_log2, grid.data())) return rc;
    if (int rc = grow(&s->d_clear, &s->cap_clear, grid.size())) return rc;
    MONTE_CUDA(cudaMemcpyAsync(s->d_clear, grid.data(), grid.size(), cudaMemcpyHostToDevice, st));
    MONTE_CUDA(cudaStreamSynchronize(st));                             // `grid` is pageable and goes out of scope
    s->cg[0] = d[0]; s->cg[1] = d[1]; s->cg[2] = d[2];
    s->clear_hash = hsh;
    memcpy(s->clear_key, key, sizeof(key));
    return MONTE_OK;
}

// (re)fill a scene: device buffers are reused when large enough; copies are issued on `st`
// labels_hash != 0: the caller has hashed `labels`; if the scene already holds exactly these bytes the 34 MB copy
// (and a clearance-grid rebuild) is skipped.  0: always copied.
// present_known: the caller's label_presence() of `labels` (one scan for all devices), or ~0u: scan here if needed
static int scene_upload(monte_mc_scene *s, const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                        const monte_mc_xs *xs, const monte_mc_spectrum *spec, cudaStream_t st, uint64_t labels_hash = 0,
                        uint32_t present_known = 0xffffffffu, size_t part_lo = 0, size_t part_hi = ~(size_t)0,
                        bool *labels_were_resident = nullptr) {
    s->geom = *g;
    McSceneDev &d = s->dev;
    const size_t nvox = (size_t)vol->nx * vol->ny * vol->nz;
    if (int rc = grow(&s->d_labels, &s->cap_labels, (nvox + 15) & ~(size_t)15)) return rc;
    const bool labels_resident = labels_hash != 0 && s->labels_hash == labels_hash && s->labels_n == nvox;
    if (labels_were_resident) *labels_were_resident = labels_resident;
    // [part_lo, part_hi): the bytes this device takes from the host (a multi-device caller completes the volume from the
    // peers afterwards, label_gather_kernel); default: everything
    if (part_hi > nvox) part_hi = nvox;
    if (!labels_resident && part_hi > part_lo)
        MONTE_CUDA(cudaMemcpyAsync((char *)s->d_labels + part_lo, labels + part_lo, part_hi - part_lo, cudaMemcpyHostToDevice, st));
    s->labels_hash = labels_hash; s->labels_n = nvox;
    d.labels = (const uint8_t *)s->d_labels;
    d.nx = vol->nx; d.ny = vol->ny; d.nz = vol->nz;
    d.inv_pitch = (float)(1.0 / vol->pitch);
    for (int a = 0; a < 3; a++) {
        d.org[a] = (float)vol->origin[a]; d.clip_lo[a] = (float)vol->clip_lo[a]; d.clip_hi[a] = (float)vol->clip_hi[a];
        d.vox_off[a] = (float)(-(double)d.org[a] * (double)d.inv_pitch - 0.5);
    }
    // per-keV tables: majorant over materials (CBCT_real325im.cu:867-868) and branching ratios (:651,656).  All small
    // tables are assembled in ONE pinned staging block of the scene and go to the device in one asynchronous copy: no
    // implicit stream synchronisation of a pageable source, so the uploads of several devices overlap.
    const int nm = xs->n_materials;
    // tracking_mode CLEARANCE needs a material to exclude; with a single material it is the reference's loop
    monte_mc_volume vres = *vol;                                   // AUTO resolved (same rule on every rank: inputs only)
    if (vol->tracking_mode == MONTE_MC_TRACK_AUTO)
        vres.tracking_mode = monte_mc_resolve_tracking(xs, spec, &vres.clearance_cell_log2, nullptr);
    const bool directional = vres.tracking_mode == MONTE_MC_TRACK_DIRECTIONAL;
    const bool adaptive = vres.tracking_mode == MONTE_MC_TRACK_ADAPTIVE || directional;
    s->heavy = vres.tracking_mode == MONTE_MC_TRACK_CLEARANCE || adaptive ? monte_xs_heavy_material(xs) : -1;
    const bool rayleigh = g->coherent_mode == MONTE_MC_COHERENT_FORMFACTOR;
    const int rn = rayleigh ? xs->ff_points : 0;
    const int n_bins = spec && spec->n_bins > 0 ? spec->n_bins : 0;
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t tab_bytes = (size_t)nm * TAB_ROWS * sizeof(float4), inv_bytes = TAB_ROWS * sizeof(float);
    const size_t cdf_bytes = n_bins ? (n_bins + 1) * sizeof(float) : 0, ray_bytes = (size_t)nm * 2 * rn * sizeof(float);
    const size_t vcs_bytes = (size_t)g->n_views * sizeof(float2);
    const size_t o_tab = 0, o_inv = al(o_tab + tab_bytes), o_invlo = al(o_inv + inv_bytes), o_vcs = al(o_invlo + 2 * inv_bytes),
                 o_ray = al(o_vcs + vcs_bytes), o_cdf = al(o_ray + ray_bytes), small_bytes = al(o_cdf + cdf_bytes) + 16;
    if (small_bytes > s->cap_small) {
        if (s->d_small) cudaFree(s->d_small);
        if (s->h_small) cudaFreeHost(s->h_small);
        s->d_small = nullptr; s->h_small = nullptr; s->cap_small = 0;
        MONTE_CUDA(cudaMalloc(&s->d_small, small_bytes));
        MONTE_CUDA(cudaHostAlloc(&s->h_small, small_bytes, cudaHostAllocPortable));
        s->cap_small = small_bytes;
    }
    char *hs = (char *)s->h_small, *ds = (char *)s->d_small;
    s->o_tab = o_tab; s->o_inv = o_inv; s->o_invlo = o_invlo; s->tables_bytes = o_vcs;
    uint32_t mask = 0xffffffffu;                                       // MAJORANT_ALL: CBCT_real325im.cu:867-868
    if (vol->majorant_mode == MONTE_MC_MAJORANT_PRESENT) {
        mask = present_known != 0xffffffffu ? present_known
             : (labels_resident && s->present_valid) ? s->present_mask : label_presence(labels, nvox, nm);
        if (!s->xs_host) s->xs_host = new monte_mc_xs();
        memcpy(s->xs_host, xs, sizeof(monte_mc_xs));
        s->present_mask = mask; s->present_valid = true; s->adaptive_host = adaptive;
    } else s->present_valid = false;
    fill_tables(s, xs, mask, adaptive);
    float2 *vcs = (float2 *)(hs + o_vcs);
    for (int v = 0; v < g->n_views; v++) {
        const double beta = M_PI * (g->angle0_deg + g->angle_step_deg * v) / 180;
        vcs[v] = make_float2((float)cos(beta), (float)sin(beta));
    }
    if (rayleigh) {                                                   // [material][x^2 grid | cumulative F^2][ff_points]
        float *ray = (float *)(hs + o_ray);
        for (int m = 0; m < nm; m++)
            for (int i = 0; i < rn; i++) { ray[((size_t)m * 2) * rn + i] = xs->ff_x2[m][i]; ray[((size_t)m * 2 + 1) * rn + i] = xs->ff_cum[m][i]; }
    }
    if (n_bins) memcpy(hs + o_cdf, spec->cdf, cdf_bytes);
    MONTE_CUDA(cudaMemcpyAsync(ds, hs, small_bytes - 16, cudaMemcpyHostToDevice, st));
    s->d_tab = ds + o_tab; s->d_inv = ds + o_inv; s->d_invlo = ds + o_invlo; s->d_view = ds + o_vcs;
    s->d_ray = rayleigh ? ds + o_ray : nullptr; s->d_cdf = n_bins ? ds + o_cdf : nullptr;
    d.tab = (const float4 *)s->d_tab; d.inv_mumax = (const float *)s->d_inv; d.n_mat = nm;
    d.n_bins = n_bins; d.cdf = (const float *)s->d_cdf; d.bin_keV = 0.5f; d.mono_keV = 140.f;
    if (spec) { d.mono_keV = (float)spec->mono_keV; d.bin_keV = (float)spec->bin_keV; }
    size_t clear_bytes = 0;
    s->vol = vres; s->n_mat_host = nm;
    if (s->heavy >= 0) {
        if (int rc = upload_clearance(s, labels, st, labels_hash)) return rc;
        clear_bytes = (size_t)s->cg[0] * s->cg[1] * s->cg[2] + 2 * TAB_ROWS * sizeof(float);
    }
    s->ray_n = rn;
    d.view_cs = (const float2 *)s->d_view;
    d.n_views = g->n_views; d.det_ny = g->ny; d.det_nx = g->nx;
    d.pixel = (float)g->pixel; d.inv_pixel = (float)(1.0 / g->pixel); d.half = (float)g->half;
    d.dso = (float)g->dso; d.dod = (float)g->dod; d.dsd = (float)(g->dso + g->dod);
    d.source_mode = g->source_mode; d.max_scatter = g->max_scatter;
    d.eid = g->detector_mode == MONTE_MC_DETECTOR_ENERGY ? 1 : 0;
    d.ring = g->detector_shape == MONTE_MC_DETECTOR_RING ? 1 : 0;
    d.ring_r = (float)g->ring_radius; d.ring_r2 = d.ring_r * d.ring_r;
    d.ring_dphi = (float)(2.0 * M_PI / g->ny); d.ring_kphi = (float)(g->ny / (2.0 * M_PI));
    if (!s->d_work) MONTE_CUDA(cudaMalloc(&s->d_work, MC_WORK_RING * sizeof(unsigned long long)));
    s->smem = (size_t)nm * TAB_ROWS * sizeof(float4) + (TAB_ROWS + 3 + d.n_bins + 1) * sizeof(float);
    s->h2d_bytes = (labels_resident || part_hi <= part_lo ? 0 : part_hi - part_lo) + tab_bytes + inv_bytes + cdf_bytes + ray_bytes + clear_bytes + vcs_bytes;
    // nothing is synchronised here: the staging block lives in the scene, and `labels` must stay valid until the
    // caller's next synchronisation of `st` (the host-buffer entry points synchronise before they return)
    return MONTE_OK;
}

int monte_gpu_scene_create(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                           const monte_mc_xs *xs, const monte_mc_spectrum *spec, monte_mc_scene **out) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_mc(g, vol, xs)) return rc;
    MONTE_ARG(labels && out, "mc: NULL argument");
    if (int rc = check_spectrum(xs, spec)) return rc;
    monte_mc_scene *s = new monte_mc_scene();
    if (int rc = scene_upload(s, g, vol, labels, xs, spec, ctx().stream)) { monte_gpu_scene_destroy(s); return rc; }
    const cudaError_t e = cudaStreamSynchronize(ctx().stream);        // `labels` may be pageable and short-lived
    if (e != cudaSuccess) { monte_gpu_scene_destroy(s); return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); }
    *out = s;
    return MONTE_OK;
}

int monte_gpu_scene_update_labels(monte_mc_scene *s, const uint8_t *labels, void *stream) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(s && labels, "scene_update_labels: NULL argument");
    const size_t nvox = (size_t)s->dev.nx * s->dev.ny * s->dev.nz;
    MONTE_CUDA(cudaMemcpyAsync(s->d_labels, labels, nvox, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    s->labels_hash = 0;
    if (s->vol.majorant_mode == MONTE_MC_MAJORANT_PRESENT && s->xs_host) {            // the majorant follows the labels
        const uint32_t mask = label_presence(labels, nvox, s->n_mat_host);
        if (mask != s->present_mask) {
            MONTE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));                 // the staging block may still be in flight
            fill_tables(s, s->xs_host, mask, s->adaptive_host);
            MONTE_CUDA(cudaMemcpyAsync(s->d_small, s->h_small, s->tables_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
            s->present_mask = mask;
        }
    }
    if (s->heavy >= 0) return upload_clearance(s, labels, (cudaStream_t)stream);   // the grid follows the labels
    return MONTE_OK;
}

void monte_gpu_scene_destroy(monte_mc_scene *s) {
    if (!s) return;
    cudaFree(s->d_labels); cudaFree(s->d_small); cudaFree(s->d_work); cudaFree(s->d_clear);
    if (s->h_small) cudaFreeHost(s->h_small);
    delete s->xs_host;
    delete s;
}

static int launch_mc(const monte_mc_scene *s, uint64_t seed, int view_begin, int view_end, uint32_t n_begin,
                     uint32_t n_end, uint32_t per, int32_t *d_image0, int32_t *d_image5, unsigned long long *d_stats,
                     uint32_t *d_fates, float *d_fate_e, cudaStream_t st) {
    MONTE_ARG(s, "mc: scene is NULL");
    MONTE_ARG(0 <= view_begin && view_begin <= view_end && view_end <= s->geom.n_views, "mc: bad view range");
    MONTE_ARG(n_begin <= n_end && n_end <= per && per > 0, "mc: bad photon range [%u,%u) of %u", n_begin, n_end, per);
    MONTE_ARG(d_image0 && d_image5, "mc: NULL image");
    const uint64_t npix = (uint64_t)s->geom.ny * s->geom.nx;
    MONTE_ARG((uint64_t)s->geom.n_views * npix * per < (1ull << 40), "mc: more than 2^40 history ids");
    MONTE_ARG(!s->dev.eid || (uint64_t)per * (MONTE_MC_TABLE_ROWS - 1) * MONTE_MC_EID_SCALE < (1ull << 31),
              "mc: %u photons per pixel overflow an int32 energy tally", per);
    // counting mode: a pixel receives at most its own `per` unscattered photons plus scattered ones from anywhere;
    // below 2^30 per pixel the int32 tallies (the reference's type, CBCT_real325im.cu:104-105) cannot wrap in
    // practice -- larger runs are split into several calls whose images the caller sums in 64 bits
    MONTE_ARG(s->dev.eid || per < (1u << 30), "mc: %u photons per pixel could overflow the int32 tallies; split the run", per);
    McLaunch L;
    L.sc = s->dev;
    {
        const uint32_t key = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u);
        for (int r = 0; r < 10; r++) L.key.rk[r] = key + (uint32_t)r * 0x9E3779B9u;
    }
    L.view_begin = view_begin; L.n_begin = n_begin; L.cnt = n_end - n_begin; L.per = per;
    L.total = (unsigned long long)(view_end - view_begin) * npix * L.cnt;
    L.n_units = (L.total + MC_UNIT - 1) / MC_UNIT;
    L.image0 = d_image0; L.image5 = d_image5; L.stats = d_stats;
    L.work = s->d_work + (s->work_next++ % MC_WORK_RING);
    L.fates = d_fates; L.fate_e = d_fate_e;
    L.ray = (const float *)s->d_ray; L.ray_n = s->ray_n;
    L.clear = (const uint8_t *)s->d_clear; L.inv_mulo = (const float *)s->d_invlo;
    L.cgx = s->cg[0]; L.cgy = s->cg[1]; L.cshift = s->cshift; L.cunit = s->cunit;
    L.clear_thr = L.inv_mulo ? L.inv_mulo + TAB_ROWS : nullptr;
    L.coct = s->coct;
    {
        static int thr = -1;                                           // MONTE_MC_SECOND=T: lanes a phase needs to run on a vote (1..32)
        if (thr < 0) { const char *e = getenv("MONTE_MC_SECOND"); thr = e ? atoi(e) : 14; if (thr < 1) thr = 1; if (thr > 32) thr = 32; }
        L.vote_bias = (uint32_t)(128 - thr) * 0x01010101u;
    }
    {
        // CBCT_real325im.cu:613-619 ends a history whose collision site lies beyond the detector plane or outside the
        // detector's extent (rotated frame).  Collision sites lie in the clip box; if the box (its circumscribed cylinder
        // about the rotation axis) lies inside those bounds for every view angle the test cannot fire
        const McSceneDev &d = s->dev;
        const float rx = fmaxf(fabsf(d.clip_lo[0]), fabsf(d.clip_hi[0])), ry = fmaxf(fabsf(d.clip_lo[1]), fabsf(d.clip_hi[1]));
        const float rz = fmaxf(fabsf(d.clip_lo[2]), fabsf(d.clip_hi[2]));
        const float rxy = sqrtf(rx * rx + ry * ry) * 1.0001f;
        L.collide_check = (d.ring ? (rxy < d.ring_r && rz < d.half) : (rxy < d.dod && rxy < d.half && rz < d.half)) ? 0u : 1u;
        L.n_vox_m1 = (uint32_t)((size_t)d.nx * d.ny * d.nz - 1);
    }
    if (L.total == 0) return MONTE_OK;
    MONTE_CUDA(cudaMemsetAsync(L.work, 0, sizeof(unsigned long long), st));
    const int sms = ctx().sm_count;
    const unsigned long long warps_needed = L.n_units;
    // K = 5 parked histories per lane is the default (measured: K=4 9.83 ms, K=5 9.45 ms, K=6 11.2 ms per
    // 1e8 C2 histories before the later trims); MONTE_MC_KERNEL=31..36 selects K=1..6 for A/B runs
    static int which_env = -1;
    if (which_env < 0) { const char *e = getenv("MONTE_MC_KERNEL"); which_env = e ? atoi(e) : 35; }
    MONTE_ARG(s->geom.n_views < 4096, "mc: more than 4095 views");
    const int rec = d_fates ? 1 : 0;
    const bool rayleigh = s->ray_n > 0;      // form-factor deflection of coherent events: its own instantiation (K = 5)
    const bool clear = s->heavy >= 0;        // two-level majorant: its own instantiations too
    const bool ring = s->dev.ring != 0;      // ring detector: its own instantiations
    const int which = rayleigh || clear || ring ? 35 : which_env;
    const int K = which >= 31 && which <= 36 ? which - 30 : (which == 44 ? 4 : (which == 43 ? 3 : 5));
    const size_t slot_bytes = (size_t)K * 32 * mc_slot_groups(rec != 0) * sizeof(uint4) * (MC_THREADS / 32) +
                              (MC_THREADS / 32) * 3 * sizeof(uint4);                      // + the per-warp source-ray cache
    L.off_inv = (uint32_t)((size_t)s->dev.n_mat * TAB_ROWS * sizeof(float4));
    L.off_cdf = L.off_inv + (TAB_ROWS + 3) * (uint32_t)sizeof(float);
    L.off_ray = L.off_cdf + (uint32_t)((s->dev.n_bins + 1 + 3) & ~3) * (uint32_t)sizeof(float);
    L.off_invlo = L.off_ray + (rayleigh ? (uint32_t)s->dev.n_mat * 2u * (uint32_t)((s->ray_n + 3) & ~3) * (uint32_t)sizeof(float) : 0u);
    L.off_slots = L.off_invlo + (clear ? 2u * (TAB_ROWS + 3) * (uint32_t)sizeof(float) : 0u);
    const size_t smem = (size_t)L.off_slots + slot_bytes;
    const void *fn = nullptr;
    switch (ring ? 210 + rec + (clear ? 2 : 0) + (rayleigh ? 4 : 0) : rayleigh || clear ? 200 + rec + (clear ? 2 : 0) + (rayleigh && clear ? 2 : 0) : which * 2 + rec) {
        case 210: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, false, false, true>; break;
        case 211: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, false, false, true>; break;
        case 212: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, false, true, true>; break;
        case 213: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, false, true, true>; break;
        case 214: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, true, false, true>; break;
        case 215: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, true, false, true>; break;
        case 216: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, true, true, true>; break;
        case 217: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, true, true, true>; break;
        case 200: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, true>; break;
        case 201: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, true>; break;
        case 202: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, false, true>; break;
        case 203: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, false, true>; break;
        case 204: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 2, true, true>; break;
        case 205: fn = (const void *)mc_transport_kernel_v3<true, 5, 3, 2, true, true>; break;
        case 62: fn = (const void *)mc_transport_kernel_v3<false, 1>; break;
        case 63: fn = (const void *)mc_transport_kernel_v3<true, 1>; break;
        case 64: fn = (const void *)mc_transport_kernel_v3<false, 2>; break;
        case 65: fn = (const void *)mc_transport_kernel_v3<true, 2>; break;
        case 66: fn = (const void *)mc_transport_kernel_v3<false, 3>; break;
        case 67: fn = (const void *)mc_transport_kernel_v3<true, 3>; break;
        case 68: fn = (const void *)mc_transport_kernel_v3<false, 4>; break;
        case 69: fn = (const void *)mc_transport_kernel_v3<true, 4>; break;
        case 70: fn = (const void *)mc_transport_kernel_v3<false, 5>; break;
        case 71: fn = (const void *)mc_transport_kernel_v3<true, 5>; break;
        case 72: fn = (const void *)mc_transport_kernel_v3<false, 6>; break;
        case 88: fn = (const void *)mc_transport_kernel_v3<false, 4, 4>; break;    // which=44: K=4, 4 CTAs/SM (64 regs)
        case 86: fn = (const void *)mc_transport_kernel_v3<false, 3, 4>; break;    // which=43
        case 73: fn = (const void *)mc_transport_kernel_v3<true, 6>; break;
        case 90: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 1>; break;   // which=45: one slot per STEP visit
        case 92: fn = (const void *)mc_transport_kernel_v3<false, 5, 3, 3>; break;   // which=46: three slots per STEP visit
        default: fn = rec ? (const void *)mc_transport_kernel_v3<true, 5> : (const void *)mc_transport_kernel_v3<false, 5>; break;
    }
    // resident CTAs per SM (persistent grid = SMs x occupancy) and the shared-memory attribute already raised, per
    // kernel variant; per-context state: forgotten at monte_gpu_shutdown (a later init may bind another device)
    struct Attr { int occ[96] = {0}; size_t smem_set[96] = {0}, smem_occ[96] = {0}; };
    static PerDev<Attr> attr_pd;                                       // function attributes are per device
    at_shutdown([] { attr_pd.get() = Attr(); });
    int *occ = attr_pd.get().occ;
    size_t *smem_set = attr_pd.get().smem_set, *smem_occ = attr_pd.get().smem_occ;
    // 94, 95 (Rayleigh), 0..3 (clearance, clearance + Rayleigh), 4..11 (ring detector): no `which` maps there (31..46 -> 62..93)
    const int slot_id = ring ? 4 + rec + (clear ? 2 : 0) + (rayleigh ? 4 : 0) : clear ? rec + (rayleigh ? 2 : 0) : rayleigh ? 94 + rec : (which * 2 + rec) % 96;
    int &oc = occ[slot_id];
    MONTE_ARG(smem <= 227 * 1024, "mc: %zu bytes of shared memory needed (> 227 KB)", smem);
    if (smem > smem_set[slot_id]) {
        MONTE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[slot_id] = smem;
    }
    if (oc == 0 || smem != smem_occ[slot_id]) {
        MONTE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, fn, MC_THREADS, smem));
        if (oc < 1) oc = 1;
        smem_occ[slot_id] = smem;
    }
    int grid = sms * oc;
    if ((unsigned long long)grid * (MC_THREADS / 32) > warps_needed) grid = (int)((warps_needed + MC_THREADS / 32 - 1) / (MC_THREADS / 32));
#ifdef MONTE_EMU   // CPU test build (tests/emu): the emulation calls the kernel function through its real type
    reinterpret_cast<void (*)(const McLaunch)>(const_cast<void *>(fn)) MONTE_CFG(grid, MC_THREADS, smem, st)(L);
#else
    void *args[] = {(void *)&L};
    MONTE_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(MC_THREADS), args, smem, st));
#endif
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

int monte_gpu_simulate_dev(const monte_mc_scene *s, uint64_t seed, int view_begin, int view_end, uint32_t n_begin,
                           uint32_t n_end, uint32_t photons_per_pixel, int32_t *d_image0, int32_t *d_image5,
                           unsigned long long *d_stats, void *stream) {
    MONTE_REQUIRE_INIT();
    return launch_mc(s, seed, view_begin, view_end, n_begin, n_end, photons_per_pixel, d_image0, d_image5, d_stats,
                     nullptr, nullptr, (cudaStream_t)stream);
}

void monte_gpu_mc_stats_unpack(const unsigned long long *w, monte_mc_stats *out) {
    if (!w || !out) return;
    out->histories = w[ST_HIST]; out->primaries = w[ST_PRIM]; out->scatter_detected = w[ST_SCAT];
    out->absorbed = w[ST_ABS]; out->interactions = w[ST_INT]; out->coherent = w[ST_COH]; out->compton = w[ST_COMP];
    out->woodcock_steps = w[ST_STEPS];
    out->sum_e_primary = (double)w[ST_EPRIM] / 1024.0;
    out->sum_e_scatter = (double)w[ST_ESCAT] / 1024.0;
}

// scene kept between monte_gpu_simulate calls, one per bound device: device buffers are reused, the tables are
// re-uploaded, the label volume only when its content hash changed
static PerDev<monte_mc_scene *> g_host_scene_pd;
#define g_host_scene (g_host_scene_pd.get())
static void mc_cleanup() {
    if (g_host_scene) monte_gpu_scene_destroy(g_host_scene);
    g_host_scene = nullptr;
}

int monte_gpu_simulate(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels, const monte_mc_xs *xs,
                       const monte_mc_spectrum *spec, uint32_t photons_per_pixel, uint64_t seed, int view_begin,
                       int view_end, int32_t *image0, int32_t *image5, monte_mc_stats *stats) {
    return monte_gpu_simulate_maps(g, vol, labels, xs, spec, photons_per_pixel, 0, photons_per_pixel, seed,
                                   view_begin, view_end, image0, image5, nullptr, nullptr, stats);
}

int monte_gpu_simulate_range(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                             const monte_mc_xs *xs, const monte_mc_spectrum *spec, uint32_t photons_per_pixel,
                             uint32_t n_begin, uint32_t n_end, uint64_t seed, int view_begin, int view_end,
                             int32_t *image0, int32_t *image5, monte_mc_stats *stats) {
    return monte_gpu_simulate_maps(g, vol, labels, xs, spec, photons_per_pixel, n_begin, n_end, seed, view_begin, view_end,
                                   image0, image5, nullptr, nullptr, stats);
}

// The whole seam of the reference's main() (CBCT_real325im.cu:232-288) in one call, on every device bound by
// monte_gpu_init: scene upload | photon range [n_begin, n_end) split evenly over the devices, one transport launch
// each | tallies of the peers summed onto device 0 and turned into -log maps in the same pass | download.
//   MONTE_MC_REDUCE=p2p  (default when all devices have peer access): tally_reduce_map_kernel, one launch on the
//                        root that loads the peers' tallies over NVLink and fuses the clamp + log epilogue
//   MONTE_MC_REDUCE=nccl: ONE ncclReduce(sum, int32) of [image0 | image5] to rank 0, then the epilogue kernel
// Both give the bits of a single-device run (integer sums; history ids are global).
int monte_gpu_simulate_maps(const monte_mc_geom *g, const monte_mc_volume *vol, const uint8_t *labels,
                            const monte_mc_xs *xs, const monte_mc_spectrum *spec, uint32_t photons_per_pixel,
                            uint32_t n_begin, uint32_t n_end, uint64_t seed, int view_begin, int view_end,
                            int32_t *image0, int32_t *image5, float *map0, float *map5, monte_mc_stats *stats) {
    MONTE_REQUIRE_INIT();
    if (int rc = check_mc(g, vol, xs)) return rc;
    MONTE_ARG(labels && image0 && image5, "mc: NULL buffer");
    if (int rc = check_spectrum(xs, spec)) return rc;
    if (view_end < 0) { MONTE_ARG(view_begin == 0, "mc: view_end < 0 (all views) needs view_begin == 0"); view_end = g->n_views; }
    MONTE_ARG(0 <= view_begin && view_begin <= view_end && view_end <= g->n_views, "mc: bad view range");
    MONTE_ARG(n_begin <= n_end && n_end <= photons_per_pixel, "mc: bad photon range [%u,%u) of %u", n_begin, n_end, photons_per_pixel);
    if (stats) memset(stats, 0, sizeof(*stats));
    if (view_begin == view_end) return MONTE_OK;                         // [v, v) is empty: nothing is written
    const int nd = n_dev();
    const char *e_red = getenv("MONTE_MC_REDUCE"), *e_cache = getenv("MONTE_MC_LABEL_CACHE");   // read per call: tests flip them
    const bool use_nccl = nd > 1 && ((e_red && !strcmp(e_red, "nccl")) || !peers_ok());
    // MONTE_MC_LABEL_CACHE unset: hash the labels only where a hit saves more than the hash costs -- a clearance grid
    // (25-110 ms of host work per new volume) or a presence scan hangs on them; for the reference's tracking loop the
    // upload itself (0.6 ms for 325^3 over PCIe, less when scattered over several devices) is cheaper than hashing
    // 34 MB on the host (0.8 ms on 8 threads), so it is simply done.  1 / 0 force either behaviour.
    int label_cache;
    if (e_cache) label_cache = atoi(e_cache);
    else {
        int tm = vol->tracking_mode, cl = 0;
        if (tm == MONTE_MC_TRACK_AUTO) tm = monte_mc_resolve_tracking(xs, spec, &cl, nullptr);
        label_cache = (tm != MONTE_MC_TRACK_GLOBAL || vol->majorant_mode == MONTE_MC_MAJORANT_PRESENT) ? 1 : 0;
    }
    const size_t nvox = (size_t)vol->nx * vol->ny * vol->nz;
    const uint64_t lhash = label_cache ? hash_labels(labels, nvox) : 0;
    uint32_t present = 0xffffffffu;                                      // majorant_mode PRESENT: one scan for all devices
    if (vol->majorant_mode == MONTE_MC_MAJORANT_PRESENT) {
        static uint64_t seen_hash = 0; static uint32_t seen_mask = 0; static int seen_nm = 0;   // (calls are not re-entrant)
        if (lhash && lhash == seen_hash && seen_nm == xs->n_materials) present = seen_mask;
        else { present = label_presence(labels, nvox, xs->n_materials); seen_hash = lhash; seen_mask = present; seen_nm = xs->n_materials; }
    }
    const size_t npix = (size_t)g->ny * g->nx, n_img = (size_t)(view_end - view_begin) * npix;
    const size_t n_cnt = 2 * n_img;
    const bool want_maps = map0 || map5;
    // per device: [image0 | image5] int32, then the stats words; on the root also the two float maps
    const size_t off_stats = (n_cnt * sizeof(int32_t) + 15) & ~(size_t)15;
    const size_t off_map = off_stats + MONTE_MC_STATS_WORDS * sizeof(unsigned long long);
    const uint32_t cnt = n_end - n_begin;
    struct Dev { int32_t *d_im = nullptr; unsigned long long *d_stats = nullptr; cudaEvent_t done = nullptr; unsigned long long w[MONTE_MC_STATS_WORDS] = {0}; } dv[MAX_DEV];
    const auto t_host0 = std::chrono::steady_clock::now();
    EventTimer *t_k[MAX_DEV] = {nullptr};
    int rc = MONTE_OK;
    // ---- scene upload.  Several devices with peer access: every device takes 1/nd of the label volume from the host and
    // the rest from its peers (MONTE_MC_LABEL_SCATTER=0: every device uploads the whole volume itself)
    const char *e_sc = getenv("MONTE_MC_LABEL_SCATTER");
    const bool scatter = nd > 1 && peers_ok() && !(e_sc && atoi(e_sc) == 0) && nvox >= (size_t)nd * 4096;
    LabelPeers lp;
    lp.n = nd;
    const unsigned long long n_words = (nvox + 15) / 16;
    for (int j = 0; j < nd; j++) lp.w_end[j] = j == nd - 1 ? n_words : n_words * (j + 1) / nd;
    cudaEvent_t lab_up[MAX_DEV] = {nullptr};
    bool need_gather = false;
    for (int i = 0; i < nd && rc == MONTE_OK; i++) {
        if ((rc = use_dev(i))) break;
        cudaStream_t st = ctx().stream;
        if (!g_host_scene) { g_host_scene = new monte_mc_scene(); at_shutdown(mc_cleanup); }
        const size_t lo = scatter ? (size_t)(i ? lp.w_end[i - 1] : 0) * 16 : 0, hi = scatter ? (size_t)lp.w_end[i] * 16 : ~(size_t)0;
        bool resident = false;
        if ((rc = scene_upload(g_host_scene, g, vol, labels, xs, spec, st, lhash, present, lo, hi, &resident))) break;
        lp.p[i] = (const uint4 *)g_host_scene->d_labels;
        if (scatter && !resident) {
            need_gather = true;
            if (cudaEventCreateWithFlags(&lab_up[i], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(lab_up[i], st) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__);
        }
    }
    for (int i = 0; i < nd && rc == MONTE_OK && need_gather; i++) {
        if ((rc = use_dev(i))) break;
        cudaStream_t st = ctx().stream;
        // (a device whose labels were resident has them complete, and nobody overwrites them: residency is decided by
        // the same hash on every device, so either all gather or none -- the events of the others are simply absent)
        for (int j = 0; j < nd && rc == MONTE_OK; j++)
            if (j != i && lab_up[j] && cudaStreamWaitEvent(st, lab_up[j], 0) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (rc || !lab_up[i]) continue;
        lp.self = i;
        label_gather_kernel MONTE_CFG((unsigned)(ctx().sm_count * 4), 256, 0, st)(lp, (uint4 *)g_host_scene->d_labels, n_words);
        if (cudaGetLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "label_gather_kernel", __FILE__, __LINE__);
    }
    for (int i = 0; i < nd && rc == MONTE_OK; i++) {
        if ((rc = use_dev(i))) break;
        Context &c = ctx();
        cudaStream_t st = c.stream;
        char *base = (char *)scratch(5, off_map + (i == 0 && want_maps ? n_cnt * sizeof(float) : 0));
        if (!base) { rc = MONTE_E_NOMEM; break; }
        dv[i].d_im = (int32_t *)base;
        dv[i].d_stats = (unsigned long long *)(base + off_stats);
        if (cudaMemsetAsync(base, 0, off_map, st) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "cudaMemsetAsync", __FILE__, __LINE__); break; }
        // photons [nb, ne) of every pixel on this device (contiguous balanced split, earlier devices take the extra)
        const uint32_t q = cnt / nd, r = cnt % nd;
        const uint32_t nb = n_begin + i * q + ((uint32_t)i < r ? i : r), ne = nb + q + ((uint32_t)i < r ? 1 : 0);
        t_k[i] = new EventTimer(st);
        t_k[i]->start();
        rc = launch_mc(g_host_scene, seed, view_begin, view_end, nb, ne, photons_per_pixel,
                       dv[i].d_im - (size_t)view_begin * npix, dv[i].d_im + n_img - (size_t)view_begin * npix, dv[i].d_stats,
                       nullptr, nullptr, st);
        t_k[i]->stop();
    }
    // ---- reduce onto device 0 + counts -> maps
    float *d_map = nullptr;
    if (rc == MONTE_OK && (rc = use_dev(0)) == MONTE_OK) {
        cudaStream_t st0 = ctx().stream;
        d_map = want_maps ? (float *)((char *)dv[0].d_im + off_map) : nullptr;
        const float log_per = logf((float)cnt);
        do {
            if (use_nccl) {
                ncclComm_t *comms = nullptr;
                if ((rc = nccl_comms(&comms))) break;
                const NcclApi *n = nccl_api();
                int r2 = n->GroupStart();
                for (int i = 0; i < nd && r2 == 0; i++)
                    r2 = n->Reduce(dv[i].d_im, dv[i].d_im, n_cnt, NCCL_INT32, NCCL_SUM, 0, comms[i], ctx_of(i).stream);
                const int r3 = n->GroupEnd();
                if (r2 || r3) { set_error("ncclReduce of the tallies failed: %s", n->GetErrorString(r2 ? r2 : r3)); rc = MONTE_E_CUDA; break; }
            } else if (nd > 1) {
                // the root's stream waits for the peers' transport kernels (events work across devices)
                for (int i = 1; i < nd; i++) {
                    use_dev(i);
                    if (cudaEventCreateWithFlags(&dv[i].done, cudaEventDisableTiming) != cudaSuccess ||
                        cudaEventRecord(dv[i].done, ctx().stream) != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "event", __FILE__, __LINE__); break; }
                }
                use_dev(0);
                for (int i = 1; i < nd && rc == MONTE_OK; i++)
                    if (cudaStreamWaitEvent(st0, dv[i].done, 0) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
                if (rc) break;
            }
            if ((nd > 1 && !use_nccl) || want_maps) {
                TallyPeers tp;
                tp.n = use_nccl ? 0 : nd - 1;
                for (int i = 0; i < tp.n; i++) tp.p[i] = dv[i + 1].d_im;
                tally_reduce_map_kernel MONTE_CFG((unsigned)((n_cnt / 4 + 1 + 255) / 256), 256, 0, st0)(tp, dv[0].d_im, n_cnt, (int32_t)cnt, log_per, d_map);
                if (cudaGetLastError() != cudaSuccess) { rc = cuda_fail(cudaGetLastError(), "tally_reduce_map_kernel", __FILE__, __LINE__); break; }
            }
        } while (0);
    }
    // ---- download (root: images and maps; every device: its counters)
    EventTimer *t_d2h = nullptr;
    if (rc == MONTE_OK) {
        cudaStream_t st0 = ctx().stream;
        t_d2h = new EventTimer(st0);
        t_d2h->start();
        cudaMemcpyAsync(image0 + (size_t)view_begin * npix, dv[0].d_im, n_img * sizeof(int32_t), cudaMemcpyDeviceToHost, st0);
        cudaMemcpyAsync(image5 + (size_t)view_begin * npix, dv[0].d_im + n_img, n_img * sizeof(int32_t), cudaMemcpyDeviceToHost, st0);
        if (map0) cudaMemcpyAsync(map0 + (size_t)view_begin * npix, d_map, n_img * sizeof(float), cudaMemcpyDeviceToHost, st0);
        if (map5) cudaMemcpyAsync(map5 + (size_t)view_begin * npix, d_map + n_img, n_img * sizeof(float), cudaMemcpyDeviceToHost, st0);
        t_d2h->stop();
        for (int i = 0; i < nd; i++) {
            use_dev(i);
            cudaMemcpyAsync(dv[i].w, dv[i].d_stats, sizeof(dv[i].w), cudaMemcpyDeviceToHost, ctx().stream);
        }
        for (int i = 0; i < nd; i++) {
            use_dev(i);
            const cudaError_t e = cudaStreamSynchronize(ctx().stream);
            if (e != cudaSuccess && rc == MONTE_OK) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
        }
    } else {
        for (int i = 0; i < nd; i++) { if (use_dev(i) == MONTE_OK) cudaStreamSynchronize(ctx().stream); }
    }
    const double ms_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    if (rc == MONTE_OK && stats) {
        unsigned long long w[MONTE_MC_STATS_WORDS] = {0};
        for (int i = 0; i < nd; i++) for (int k = 0; k < MONTE_MC_STATS_WORDS; k++) w[k] += dv[i].w[k];
        monte_gpu_mc_stats_unpack(w, stats);
        for (int i = 0; i < nd; i++) if (t_k[i]) { use_dev(i); stats->ms_kernel = fmax(stats->ms_kernel, t_k[i]->ms()); }
        use_dev(0);
        stats->ms_d2h = t_d2h ? t_d2h->ms() : 0;
        stats->ms_total = ms_host;
        stats->ms_h2d = fmax(0.0, ms_host - stats->ms_kernel - stats->ms_d2h);   // scene upload + hashing (host clock)
        stats->launches = nd + ((nd > 1 && !use_nccl) || want_maps ? 1 : 0); stats->sm_count = ctx_of(0).sm_count;
    }
    for (int i = 0; i < nd; i++) {
        if (use_dev(i) != MONTE_OK) continue;
        delete t_k[i];
        if (dv[i].done) cudaEventDestroy(dv[i].done);
        if (lab_up[i]) cudaEventDestroy(lab_up[i]);
        if (rc != MONTE_OK && g_host_scene) g_host_scene->labels_hash = 0;    // (a scattered upload may be incomplete)
    }
    use_dev(0);
    delete t_d2h;
    return rc;
}

int monte_gpu_simulate_fates(const monte_mc_scene *s, uint64_t seed, int view, uint32_t photons_per_pixel,
                             uint32_t *fates, float *energies) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(s && fates, "mc_fates: NULL argument");
    MONTE_ARG(view >= 0 && view < s->geom.n_views, "mc_fates: bad view");
    const size_t npix = (size_t)s->geom.ny * s->geom.nx, nh = npix * photons_per_pixel;
    MONTE_ARG(nh < (1ull << 28), "mc_fates: too many histories for a fate dump");
    cudaStream_t st = ctx().stream;
    const size_t n_img = (size_t)s->geom.n_views * npix;
    char *base = (char *)scratch(6, 2 * n_img * sizeof(int32_t) + nh * 8);
    if (!base) return MONTE_E_NOMEM;
    int32_t *d_im = (int32_t *)base;
    uint32_t *d_f = (uint32_t *)(base + 2 * n_img * sizeof(int32_t));
    float *d_e = (float *)(d_f + nh);
    MONTE_CUDA(cudaMemsetAsync(base, 0, 2 * n_img * sizeof(int32_t) + nh * 8, st));
    if (int rc = launch_mc(s, seed, view, view + 1, 0, photons_per_pixel, photons_per_pixel, d_im, d_im + n_img, nullptr, d_f, d_e, st)) return rc;
    MONTE_CUDA(cudaMemcpyAsync(fates, d_f, nh * 4, cudaMemcpyDeviceToHost, st));
    if (energies) MONTE_CUDA(cudaMemcpyAsync(energies, d_e, nh * 4, cudaMemcpyDeviceToHost, st));
    MONTE_CUDA(cudaStreamSynchronize(st));
    return MONTE_OK;
}

int monte_gpu_counts_to_map_dev(const int32_t *d_counts, size_t n, int32_t per, float *d_map, void *stream) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(d_counts && d_map && per > 0, "counts_to_map: bad argument");
    if (n == 0) return MONTE_OK;
    counts_to_map_kernel MONTE_CFG((unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream)(d_counts, n, per, logf((float)per), d_map);
    MONTE_CUDA(cudaGetLastError());
    return MONTE_OK;
}

int monte_gpu_counts_to_map(const int32_t *counts, size_t n, int32_t per, float *map) {
    MONTE_REQUIRE_INIT();
    MONTE_ARG(counts && map && per > 0, "counts_to_map: bad argument");
    if (n == 0) return MONTE_OK;
    cudaStream_t st = ctx().stream;
    char *base = (char *)scratch(7, n * 8);
    if (!base) return MONTE_E_NOMEM;
    MONTE_CUDA(cudaMemcpyAsync(base, counts, n * 4, cudaMemcpyHostToDevice, st));
    if (int rc = monte_gpu_counts_to_map_dev((const int32_t *)base, n, per, (float *)(base + n * 4), st)) return rc;
    MONTE_CUDA(cudaMemcpyAsync(map, base + n * 4, n * 4, cudaMemcpyDeviceToHost, st));
    MONTE_CUDA(cudaStreamSynchronize(st));
    return MONTE_OK;
}

}  // extern "C"
