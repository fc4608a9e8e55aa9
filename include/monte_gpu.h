/* libmonte_gpu — C ABI of the B200-native CBCT engine (MC photon transport + FDK).
 *
 * The reference (Tomato27/Monte) has no plugin/FFI layer: every program is one
 * main().  The seams this ABI replaces are
 *   - FDK:  recon/bp3d20.cpp:29 (fread of the map) .. :171 (first writeRawFile);
 *           same span in recon/bp3d20_325.cpp:30..:181 and recon/fbp2.cpp:28..:157
 *   - MC:   monte_cu/CBCT_real325im.cu:232-248 (RandStateGenerator + projection
 *           launch + the two D2H copies) and :258-288 (counts -> -log map);
 *           CPU form monte_cpp/CBCT_real2.cpp:169-595 (the view/pixel/photon loops)
 *   - tables: monte_cpp/CBCT_real2.cpp:633-668 (readcsv), CBCT_real325im.cu:299-355
 * Plain pointers and sizes only.  All host buffers are caller-owned; outputs are
 * overwritten; inputs are const.  Every call returns 0 on success or a negative
 * MONTE_E_* code (never exit()), with text in monte_gpu_last_error().
 * Calls are blocking and not re-entrant (reference: single host thread,
 * cudaThreadSynchronize after each launch, CBCT_real325im.cu:236,244).
 *
 * Lengths are cm, energies keV, angles degrees — the reference's units.
 */
#ifndef MONTE_GPU_H
#define MONTE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MONTE_GPU_ABI_VERSION 7   /* 2: monte_mc_geom.detector_mode; 3: monte_mc_geom.coherent_mode (was
                                     `reserved`, 0 = unchanged behaviour), form-factor tables appended to monte_mc_xs,
                                     tracking_mode / clearance_cell_log2 appended to monte_mc_volume (0 = unchanged);
                                     4: monte_gpu_init binds 1..8 devices and the host-buffer calls monte_gpu_simulate* /
                                     monte_gpu_fdk shard over them; monte_gpu_simulate_maps; "all views" is spelled
                                     view_end < 0 (the range [0, 0) is now empty, as in the device forms);
                                     5: majorant_mode appended to monte_mc_volume (0 = unchanged); monte_hu_class,
                                     monte_hu_classes_default, monte_ctnum_segment (N-class HU segmentation);
                                     6: detector_shape / ring_radius appended to monte_mc_geom (0 = the flat panel);
                                     7: monte_gpu_ipc_export / _open / _close, monte_gpu_fdk_backproject_peers_dev       */

/* ---- status codes ------------------------------------------------------ */
#define MONTE_OK            0
#define MONTE_E_ARG        -1   /* bad argument / inconsistent geometry          */
#define MONTE_E_CUDA       -2   /* CUDA runtime error (text in last_error)       */
#define MONTE_E_NOINIT     -3   /* monte_gpu_init not called                     */
#define MONTE_E_NODEV      -4   /* no usable sm_100 device                        */
#define MONTE_E_NOMEM      -5
#define MONTE_E_IO         -6   /* table / raw file problems (host helpers)      */

/* ---- library lifetime --------------------------------------------------- */
/* Binds ndev (1..8) devices of one NVSwitch box to this process; ids[] are CUDA ordinals
 * (ids == NULL -> 0..ndev-1).  With ndev == 1 every call runs on that device (one process
 * per GPU, the form the torch.distributed plumbing uses).  With ndev > 1 (one
 * stream per device, peer access over NVLink, an in-library NCCL communicator created on
 * first use) the host-buffer calls monte_gpu_simulate / _range / _maps split the photon
 * range over the devices and sum the tallies onto ids[0], and monte_gpu_fdk shards the
 * filter by views and the backprojection by z-slabs of equal work; results are bit-identical
 * to ndev == 1.  All other calls (device-pointer forms, scenes, projector, fbp2) run on ids[0]. */
int  monte_gpu_init(int ndev, const int *ids);
int  monte_gpu_device_count(void);     /* devices bound by the last monte_gpu_init (0 before)     */
int  monte_gpu_peer_access(void);      /* 1: every bound device can address every other (NVLink P2P) */
void monte_gpu_shutdown(void);
const char *monte_gpu_last_error(void);
int  monte_gpu_abi_version(void);
/* number of SMs of the bound device (grid sizing is reported in stats) */
int  monte_gpu_sm_count(void);

/* ======================================================================== */
/*  FDK reconstruction  (replaces bp3d20 / bp3d20_325 / fbp2)               */
/* ======================================================================== */

/* weight_mode */
#define MONTE_FDK_REFERENCE 0  /* reproduces the shipped arithmetic incl. its quirks:
                                  cosine weight and distance weight use weight_dist
                                  (=Dod=60, bp3d20.cpp:40,159), tan() of the angle in
                                  degrees taken as radians (bp3d20.cpp:137), fixed
                                  filter_scale (bp3d20.cpp:68), out_scale fudge (:160) */
#define MONTE_FDK_TEXTBOOK  1  /* Feldkamp 1984: weights with Dsd / Dso, ramp scaled by
                                  the detector pitch, no fudge factors                 */

/* coord_mode: association of the detector-coordinate expression (1-ulp effects) */
#define MONTE_FDK_COORD_SCALE_AFTER  0  /* y = -inv_du*(u - half_u)   bp3d20.cpp:123-124     */
#define MONTE_FDK_COORD_SCALE_BEFORE 1  /* y = -(u*inv_du - nu/2)     bp3d20_325.cpp:134-135 */

typedef struct monte_fdk_geom {
    /* projections: map[view][iu][iv], iu = transaxial (the filtered direction,
       "zeta"/"c" in bp3d20.cpp:37,67), iv = axial ("p"/"d").  Filtered output is
       written transposed, filtered[view][iv][iu] (bp3d20.cpp:70).                 */
    int32_t n_views;
    int32_t nu, nv;
    double  du, dv;            /* detector pitch: 0.5 (bp3d20) or 0.1 (bp3d20_325)   */
    double  half_u, half_v;    /* 16.25: detector half-extent (bp3d20.cpp:40,116)    */
    double  dso, dsd;          /* 160, 220  (bp3d20.cpp:85,107)                      */
    double  weight_dist;       /* 60 in REFERENCE mode (Q10); ignored in TEXTBOOK    */
    double  filter_scale;      /* 0.5 (bp3d20.cpp:68)                                */
    double  out_scale;         /* 2.7 (bp3d20.cpp:160)                               */
    double  out_scale2;        /* 1, or 5 for bp3d20_325.cpp:170 ((o*2.7)*5)         */
    double  angle0_deg;        /* first view angle                                   */
    double  angle_step_deg;    /* "beta_span" = 1 (bp3d20.cpp:82)                    */
    /* volume: vol_xy[z][t][s], X = x0 + vox*s, Y = y0 - vox*t, Z = z0 - vox*z
       (bp3d20.cpp:96-98: -12.8+0.1s, 12.8-0.1t, 12.8-0.1z)                          */
    int32_t nx, ny, nz;
    double  vox;
    double  x0, y0, z0;
    /* region actually reconstructed (everything else stays 0); the shipped loops
       run s in [125,130) only (bp3d20.cpp:93).                                      */
    int32_t s_begin, s_end, t_begin, t_end, z_begin, z_end;
    /* sphere mask (bp3d20.cpp:145): (z-cz)^2+(t-ct)^2+(s-cs)^2 <= r2 ; r2 < 0 = off  */
    int32_t mask_cs, mask_ct, mask_cz;
    int64_t mask_r2;
    int32_t weight_mode;       /* MONTE_FDK_REFERENCE / MONTE_FDK_TEXTBOOK           */
    int32_t coord_mode;        /* MONTE_FDK_COORD_*                                  */
} monte_fdk_geom;

typedef struct monte_fdk_stats {
    double   ms_h2d, ms_filter, ms_backproject, ms_transpose, ms_d2h, ms_total;
    uint64_t voxel_updates;    /* voxels in ROI x views                              */
    uint64_t filter_macs;
    int32_t  launches;         /* kernels launched by this call                      */
    int32_t  sm_count;
} monte_fdk_stats;

/* Presets filling every field with the literals of the shipped programs. */
void monte_fdk_geom_bp3d20(monte_fdk_geom *g);      /* recon/bp3d20.cpp        */
void monte_fdk_geom_bp3d20_325(monte_fdk_geom *g);  /* recon/bp3d20_325.cpp    */

/* Whole pipeline on host buffers (the drop-in for bp3d20.cpp:29-171).  With several bound devices (monte_gpu_init):
 * device i uploads and filters views [n_views i/N, n_views (i+1)/N), then backprojects all views into z-slab i of
 * monte_gpu_fdk_partition, loading the detector rows that slab reads straight out of its peers' memory over NVLink
 * (the exchange is fused into the backprojector's row-pair conversion), and copies the slab into vol_xy in place.
 * map      [n_views][nu][nv]  float32, in
 * filtered [n_views][nv][nu]  float32, out, nullable   ("map_out", bp3d20.cpp:187)
 * vol_xy   [nz][ny][nx]       float32, out             ("image_xy")
 * vol_zy   [nx][ny][nz]       float32, out, nullable   ("image_zy", bp3d20.cpp:161)  */
int monte_gpu_fdk(const monte_fdk_geom *g, const float *map, float *filtered,
                  float *vol_xy, float *vol_zy, monte_fdk_stats *stats);

/* Device-resident stages (pointers are CUDA device pointers on the bound device,
 * stream is a cudaStream_t passed as void*; asynchronous w.r.t. the host).
 * The filtered projections live in a padded row layout:
 *   rows = n_views*nv + 2, row pitch = monte_gpu_fdk_filtered_pitch(g) floats,
 *   element [r][nu] duplicates [r+1][0] and the two trailing rows are zero, so the
 *   reference's inclusive bilinear bound (bp3d20.cpp:152-156, reads one past the
 *   row / view) is reproduced by plain row addressing.                             */
size_t monte_gpu_fdk_filtered_pitch(const monte_fdk_geom *g);           /* floats */
size_t monte_gpu_fdk_filtered_elems(const monte_fdk_geom *g);           /* floats */
/* Steps 1+2 of bp3d20.cpp (:36-43 cosine weight, :48-73 Ram-Lak convolution, written transposed) for
 * views [view_begin, view_end).  nu <= 256: direct shared-memory convolution in the reference's
 * summation order; wider detectors (up to nu = 2048): the same linear convolution by FFT (length >= 2 nu,
 * fp32, ~3e-7 of the row maximum from the direct sum).  MONTE_FDK_FILTER=direct|fft overrides.        */
int monte_gpu_fdk_filter_dev(const monte_fdk_geom *g, const float *d_map,
                             int view_begin, int view_end,
                             float *d_filtered_padded, void *stream);
/* fix up the duplicated column / trailing rows after views were written or gathered */
int monte_gpu_fdk_pad_dev(const monte_fdk_geom *g, float *d_filtered_padded, void *stream);
/* Axial detector rows [row_lo, row_hi) of every view that backprojecting slices [z_lo, z_hi) reads, beside
 * rows 0..3 of every view (what the previous view's last row reaches into, bp3d20.cpp:152-156).  A rank that
 * owns a z-slab needs only these rows of the other ranks' filtered views (multi-GPU exchange).
 * row_lo == row_hi: no view sees the slab.                                                               */
int monte_gpu_fdk_slab_rows(const monte_fdk_geom *g, int z_lo, int z_hi, int *row_lo, int *row_hi);
/* z-slab cuts of equal modelled work, z_cuts[0] = 0 <= ... <= z_cuts[n_parts] = nz (on multiples of 16 slices where
 * possible): at wide cone angles the end slices see the detector in few views or none (bp3d20.cpp:116 skips them), so
 * slabs of equal thickness would leave the end devices idle.  This is the partition monte_gpu_fdk uses with n_parts =
 * bound devices; any partition gives the same voxels bit for bit.  Host only.                                       */
int monte_gpu_fdk_partition(const monte_fdk_geom *g, int n_parts, int *z_cuts);
/* Backproject all views into z-slices [z_lo, z_hi) of the volume;
 * d_vol_slab points at slice z_lo, layout [z_hi-z_lo][ny][nx].  Overwrites.        */
int monte_gpu_fdk_backproject_dev(const monte_fdk_geom *g, const float *d_filtered_padded,
                                  int z_lo, int z_hi, float *d_vol_slab, void *stream);
/* Same for views [view_lo, view_hi) only.  continue_sum != 0: the slab already holds the fp32 partial
 * sums of the earlier views (they are continued exactly, so feeding the views in ascending pieces
 * gives the bits of the single call); 0: the slab is overwritten.  This is what lets a multi-GPU
 * host overlap the exchange of later views with the backprojection of earlier ones.              */
int monte_gpu_fdk_backproject_views_dev(const monte_fdk_geom *g, const float *d_filtered_padded,
                                        int z_lo, int z_hi, float *d_vol_slab,
                                        int view_lo, int view_hi, int continue_sum, void *stream);
/* pad fix-up restricted to the rows of views [view_lo, view_hi) (+ the two rows they reach into) */
int monte_gpu_fdk_pad_views_dev(const monte_fdk_geom *g, float *d_filtered_padded,
                                int view_lo, int view_hi, void *stream);
/* One process per GPU: backproject all views into slices [z_lo, z_hi) with the filtered rows left where they
 * were computed -- n_seg segments of views in ascending order, segment o = views [seg_v_end[o-1], seg_v_end[o]) whose
 * padded-row layout (monte_gpu_fdk_filter_dev output, pad not needed) starts at seg_base[o] such that row R of the GLOBAL
 * layout lives at seg_base[o] + R * pitch.  seg_base[o] may point into a peer process's memory opened with
 * monte_gpu_ipc_open: the detector-row band the slab reads is loaded over NVLink while it is converted to the
 * backprojector's pair layout (the kernel the in-library multi-device call uses), so the exchange needs no collective,
 * no packing and no second copy.  The caller orders the processes (the peers' filters must have finished; nobody may
 * overwrite its rows before every peer is done).  n_seg <= 32.                                                     */
int monte_gpu_fdk_backproject_peers_dev(const monte_fdk_geom *g, int n_seg, const float *const *seg_base, const int *seg_v_end,
                                        int z_lo, int z_hi, float *d_vol_slab, void *stream);
/* CUDA IPC plumbing for such hosts.  export: a 64-byte handle of the device allocation that contains d_ptr and d_ptr's
 * byte offset in it (send both to the peer process).  open: map that allocation into this process (peer access is
 * enabled lazily) and return the address of the same byte.  close: unmap (pass the pointer open returned).          */
#define MONTE_IPC_HANDLE_BYTES 64
int monte_gpu_ipc_export(const void *d_ptr, unsigned char handle[MONTE_IPC_HANDLE_BYTES], uint64_t *offset);
int monte_gpu_ipc_open(const unsigned char handle[MONTE_IPC_HANDLE_BYTES], uint64_t offset, void **d_ptr);
int monte_gpu_ipc_close(void *d_ptr);
/* vol_zy[s][t][z] = vol_xy[z][t][s] */
int monte_gpu_fdk_transpose_dev(const monte_fdk_geom *g, const float *d_vol_xy,
                                float *d_vol_zy, void *stream);
/* copy padded -> dense [n_views][nv][nu] (device to device) */
int monte_gpu_fdk_unpad_dev(const monte_fdk_geom *g, const float *d_filtered_padded,
                            float *d_filtered_dense, void *stream);

/* 2-D fan-beam FBP (recon/fbp2.cpp): sino[n_views][nu] -> filtered[n_views][nu],
 * image[ny][nx]; nearest-neighbour detector lookup (fbp2.cpp:126,147), views start
 * at view_first (=1 in the shipped loop, fbp2.cpp:89), out_scale 1.7 (:148).
 * Uses g->nu, du, half_u, dso, dsd, weight_dist, filter_scale, out_scale, angles,
 * nx, ny, vox, x0, y0, s/t ROI.                                                    */
void monte_fdk_geom_fbp2(monte_fdk_geom *g);
int monte_gpu_fbp2(const monte_fdk_geom *g, int view_first, const float *sino,
                   float *filtered, float *image, monte_fdk_stats *stats);

/* ======================================================================== */
/*  Monte-Carlo photon transport (replaces the `projection` kernel)         */
/* ======================================================================== */

#define MONTE_MC_MAX_MATERIALS 8
#define MONTE_MC_TABLE_ROWS    201   /* index = keV, 0..200 (CBCT_real325im.cu:76-78) */
#define MONTE_MC_FF_POINTS     128   /* grid points of a Rayleigh form-factor table */

/* cross-section tables per material, index (int)(E+0.5) (CBCT_real325im.cu:501,627-630).
 * Values are mass coefficients cm^2/g; mu = value * density.                        */
typedef struct monte_mc_xs {
    int32_t n_materials;                 /* label m (1..n) uses entry m-1; label 0 = air;
                                            labels > n use the last entry (the reference's
                                            final else = PMMA, CBCT_real325im.cu:640-646) */
    float   density[MONTE_MC_MAX_MATERIALS];                       /* g/cm^3 */
    float   coh  [MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];    /* coherent           */
    float   compt[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];    /* incoherent         */
    float   photo[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];    /* photoelectric "ab" */
    float   total[MONTE_MC_MAX_MATERIALS][MONTE_MC_TABLE_ROWS];    /* "mua"              */
    /* Rayleigh form-factor tables, read only when monte_mc_geom.coherent_mode == MONTE_MC_COHERENT_FORMFACTOR
     * (the reference has none: its coherent event keeps the direction, CBCT_real325im.cu:656-695).
     * Per material: ff_x2[i] = x^2 grid, x = sin(theta/2)/lambda in 1/Angstrom, ascending from ff_x2[0] = 0 and
     * reaching at least (E_max/12.398 keV A)^2; ff_cum[i] = integral_0^{x2_i} F(x)^2 d(x^2), any normalisation.
     * monte_xs_formfactor_hydrogenic() fills them with an analytic stand-in.                                   */
    int32_t ff_points;                                             /* 0 = no tables; else 2..MONTE_MC_FF_POINTS */
    float   ff_x2 [MONTE_MC_MAX_MATERIALS][MONTE_MC_FF_POINTS];
    float   ff_cum[MONTE_MC_MAX_MATERIALS][MONTE_MC_FF_POINTS];
} monte_mc_xs;

/* voxelised label volume, x fastest (make_image01.cpp:20: g[k*L*M + j*M + i]).
 * Voxel index along an axis = (int)floor((p - origin)/pitch); the reference's two
 * addressing forms map onto it: CBCT_real325im.cu:921 int((p+10)*10) -> origin -10;
 * CBCT_real2.cpp:771 rint(p*10)+90 -> origin -(90+0.5)*0.1.
 * Lookups happen only inside the clip box (CBCT_real2.cpp:770; the tight box around
 * the phantom); outside it the photon flies straight (air).  A position exactly on a voxel
 * face may be assigned to either neighbour (the kernel rounds to even, the CPU restatement
 * floors): a null set, except for pencil rays that run along a face of an even-sized volume
 * centred on the axis -- use odd sizes, as the reference's 325^3 / 185x185x325 volumes are.  */
/* tracking_mode: how tentative collisions are sampled.  Both are exact (same physics, same expected tallies);
 * they consume different variates, so a run is reproducible bit for bit only within one mode.               */
#define MONTE_MC_TRACK_GLOBAL    0  /* the reference's Woodcock loop: one majorant, the maximum over all materials
                                       at the photon's energy (CBCT_real325im.cu:866-868, :886-968)                 */
#define MONTE_MC_TRACK_CLEARANCE 1  /* two-level majorant: a coarse clearance grid (monte_mc_clearance_grid) gives
                                       every cell a radius D inside which the heaviest material does not occur; a
                                       flight that starts there is sampled with the majorant of the other materials
                                       and, if it would go farther than D, stops at D without a collision (Woodcock
                                       steps are memoryless).  Pays when a dense insert sets the global majorant far
                                       above the bulk: polyenergetic spectra, calcium / bone in water               */

#define MONTE_MC_TRACK_ADAPTIVE  3  /* CLEARANCE with a per-step test: the light majorant is used only where the clearance
                                       D exceeds -ln(1 - mu_light/mu_max)/mu_light at the photon's energy, i.e. where
                                       stopping at D is less likely than a virtual collision would be; elsewhere the
                                       step is the reference's.  Never more tentative collisions than GLOBAL (oracle:
                                       4.32 vs 4.47 per history at 140 keV, 4.93 vs 22.4 at 120 kVp)                */
#define MONTE_MC_TRACK_DIRECTIONAL 4 /* ADAPTIVE with one clearance per octant of the flight direction
                                       (monte_mc_clearance_grid_octants): only the heavy material a ray with that direction
                                       can still reach counts, so a photon flying away from the insert is never cut short.
                                       Oracle, steps per history: 3.6 at 140 keV (reference loop 4.5), 3.6 at 120 kVp (22.4) */
#define MONTE_MC_TRACK_AUTO      2  /* the library picks from the tables and the spectrum (monte_mc_resolve_tracking):
                                       DIRECTIONAL with 4-voxel cells if the global majorant is on average more than
                                       3x the majorant of the lighter materials, else GLOBAL                        */

/* majorant_mode: which materials the per-keV Woodcock majorant is taken over.  Exact either way (a majorant only has
 * to bound the attenuation that occurs); a run is reproducible bit for bit only within one mode.               */
#define MONTE_MC_MAJORANT_ALL     0  /* every table that was loaded (CBCT_real325im.cu:866-868)                    */
#define MONTE_MC_MAJORANT_PRESENT 1  /* only the materials that occur in the label volume (monte_xs_majorant): a
                                        segmented CT volume without dense bone is not tracked with bone's majorant.
                                        The library scans the labels when they are uploaded (8 host threads)       */

typedef struct monte_mc_volume {
    int32_t nx, ny, nz;
    double  pitch;
    double  origin[3];
    double  clip_lo[3], clip_hi[3];
    int32_t tracking_mode;        /* MONTE_MC_TRACK_*; 0 = the reference's single majorant                      */
    int32_t clearance_cell_log2;  /* CLEARANCE / ADAPTIVE: cells of 2^n voxels per side (0..8); 2 or 3         */
    int32_t majorant_mode;        /* MONTE_MC_MAJORANT_*; 0 = the reference's maximum over all tables           */
    int32_t reserved0;            /* 0                                                                          */
} monte_mc_volume;

#define MONTE_MC_SOURCE_PENCIL 0  /* one pencil per pixel centre, `per` photons each
                                     (CBCT_real325im.cu:464-527; survey Q2)             */
#define MONTE_MC_SOURCE_CONE   1  /* same stratification, aim point jittered uniformly
                                     inside the pixel: a sampled cone beam              */

/* detector response (anything but COUNTING changes results against the reference) */
#define MONTE_MC_DETECTOR_COUNTING 0  /* image[...]++ per detected photon (CBCT_real325im.cu:589-590)   */
#define MONTE_MC_DETECTOR_ENERGY   1  /* energy integrating: += (int)(E_keV * MONTE_MC_EID_SCALE + 0.5) */
#define MONTE_MC_EID_SCALE 16         /* tally unit = 1/16 keV; per * 200 keV * 16 must stay < 2^31     */

/* coherent (Rayleigh) events (anything but FORWARD changes results against the reference) */
#define MONTE_MC_COHERENT_FORWARD    0  /* no deflection: the photon flies on (CBCT_real325im.cu:656-695)       */
#define MONTE_MC_COHERENT_FORMFACTOR 1  /* theta from (1+cos^2 theta)/2 * F(x)^2 with the tables in monte_mc_xs:
                                           x^2 drawn from F^2 up to x^2_max = (E/12.398)^2, accepted with
                                           probability (1+cos^2 theta)/2, cos theta = 1 - 2 x^2/x^2_max; phi
                                           uniform; energy unchanged                                             */

/* detector shape.  RING is the geometry of the reference's 2-D programs (monte_cpp/circle3_2.cpp:
 * a source at the centre of a water disc, a ring of 180 angular bins at radius 10 cm, :162,:252), generalised to a
 * cylinder: the source sits on the rotation axis at the origin, bin (i, j) of the [ny][nx] images covers the angle
 * [2 pi i / ny, 2 pi (i + 1) / ny) about the z axis (rotated by the view angle) and the height
 * [half - pixel (j + 1), half - pixel j] at radius ring_radius; nx = 1 with max_scatter as wanted is the 2-D case.
 * `per` photons are aimed at the centre of every bin (SOURCE_PENCIL) or uniformly inside it (SOURCE_CONE); dso / dod
 * are not used.  The clip box must lie inside the ring.                                                          */
#define MONTE_MC_DETECTOR_FLAT 0      /* the flat panel at x = dod (CBCT_real325im.cu:459-466, :567-590)          */
#define MONTE_MC_DETECTOR_RING 1

typedef struct monte_mc_geom {
    int32_t n_views;
    double  angle0_deg, angle_step_deg;  /* view v at angle0 + v*step (1 deg, :462,508) */
    int32_t ny, nx;          /* detector pixels: ny transaxial ("i"), nx axial ("j")    */
    double  pixel;           /* 0.1 (GPU) / 0.5 (CPU)                                    */
    double  half;            /* detector_height 16.25 (CBCT_real325im.cu:459).  Pixel i covers [half - pixel*(i+1),
                                half - pixel*i] on both axes; the reference's detector is square and centred,
                                n * pixel = 2 * half (its bin formula :574-575 assumes that)              */
    double  dso, dod;        /* 160, 60                                                  */
    int32_t source_mode;
    int32_t max_scatter;     /* ScatterNUM = 5 (CBCT_real325im.cu:7)                     */
    int32_t detector_mode;   /* MONTE_MC_DETECTOR_*; 0 = the reference's photon counting  */
    int32_t coherent_mode;   /* MONTE_MC_COHERENT_*; 0 = the reference's undeflected coherent event */
    int32_t detector_shape;  /* MONTE_MC_DETECTOR_FLAT / _RING; 0 = the reference's flat panel               */
    int32_t reserved1;       /* 0                                                                         */
    double  ring_radius;     /* RING only: radius of the detector cylinder, cm                            */
} monte_mc_geom;

/* spectrum: n_bins == 0 -> mono-energetic at mono_keV (as shipped: 140, survey Q3);
 * else cdf[0..n_bins], cdf[0]=0, cdf[n_bins]=1, E = (k+1)*bin_keV for
 * cdf[k] <= u <= cdf[k+1]  (CBCT_real325im.cu:493-497, bin 0.5 keV).                   */
typedef struct monte_mc_spectrum {
    int32_t n_bins;
    double  bin_keV;
    double  mono_keV;
    const float *cdf;
} monte_mc_spectrum;

typedef struct monte_mc_stats {
    uint64_t histories;
    uint64_t primaries;        /* histories that reached the detector unscattered        */
    uint64_t scatter_detected; /* tallied to image5 after >=1 interaction                */
    uint64_t absorbed;         /* photoelectric                                          */
    uint64_t interactions;     /* all accepted collisions                                */
    uint64_t coherent, compton;
    uint64_t woodcock_steps;   /* tentative collisions sampled (N-bar * histories)       */
    double   sum_e_primary;    /* keV, for mean detected energy                          */
    double   sum_e_scatter;
    double   ms_h2d, ms_kernel, ms_d2h, ms_total;
    int32_t  launches;
    int32_t  sm_count;
} monte_mc_stats;

/* Whole simulation on host buffers.  photons_per_pixel is the reference's `per`
 * (CBCT_real325im.cu:182): histories per view = per*ny*nx.  History id
 * h = (view*ny*nx + pixel)*per + n keys a Philox2x32-10 counter (seed, h), so the
 * result is independent of how histories are partitioned (lanes, CTAs, devices).  view_end < 0
 * (with view_begin == 0) means all views; [v, v) is empty; only the requested views of
 * image0/image5 are written.
 * image0 [n_views][ny][nx] int32: unscattered;  image5: unscattered + scattered
 * (CBCT_real325im.cu:584-585,692,841).  labels: uint8 [nz][ny][nx].                    */
int monte_gpu_simulate(const monte_mc_geom *g, const monte_mc_volume *vol,
                       const uint8_t *labels, const monte_mc_xs *xs,
                       const monte_mc_spectrum *spec, uint32_t photons_per_pixel,
                       uint64_t seed, int view_begin, int view_end,
                       int32_t *image0, int32_t *image5, monte_mc_stats *stats);

/* Same, restricted to photons n in [n_begin, n_end) of every pixel: several processes (one per
 * GPU) split a view and sum their integer tallies; the union equals the undivided run.          */
int monte_gpu_simulate_range(const monte_mc_geom *g, const monte_mc_volume *vol,
                             const uint8_t *labels, const monte_mc_xs *xs,
                             const monte_mc_spectrum *spec, uint32_t photons_per_pixel,
                             uint32_t n_begin, uint32_t n_end,
                             uint64_t seed, int view_begin, int view_end,
                             int32_t *image0, int32_t *image5, monte_mc_stats *stats);

/* The whole seam of the reference's main() in one call (CBCT_real325im.cu:232-288: launch, D2H of both count
 * images, clamp + -log maps): as monte_gpu_simulate_range, and additionally map0 / map5 (nullable, float32
 * [n_views][ny][nx]) = -ln(clamp(image, 1, I0)) + ln(I0), I0 = n_end - n_begin, computed on the device in the same
 * pass that sums the per-device tallies (fused reduce + epilogue).  With several bound
 * devices: MONTE_MC_REDUCE=p2p (default; the root kernel loads the peers' tallies over NVLink) or =nccl (one
 * ncclReduce to ids[0], then the epilogue kernel).  The label volume is re-uploaded only when its content changed
 * (64-bit hash of the host buffer) wherever a clearance grid or a presence scan hangs on it; for the reference's tracking
 * loop uploading is cheaper than hashing and is simply done (MONTE_MC_LABEL_CACHE=1 / 0 force either behaviour).        */
int monte_gpu_simulate_maps(const monte_mc_geom *g, const monte_mc_volume *vol,
                            const uint8_t *labels, const monte_mc_xs *xs,
                            const monte_mc_spectrum *spec, uint32_t photons_per_pixel,
                            uint32_t n_begin, uint32_t n_end,
                            uint64_t seed, int view_begin, int view_end,
                            int32_t *image0, int32_t *image5, float *map0, float *map5,
                            monte_mc_stats *stats);

/* Device-resident form: the scene is uploaded once, tallies accumulate into device
 * int32 images [view_end-view_begin... indexed by absolute view][ny][nx].             */
typedef struct monte_mc_scene monte_mc_scene;   /* opaque */
int monte_gpu_scene_create(const monte_mc_geom *g, const monte_mc_volume *vol,
                           const uint8_t *labels, const monte_mc_xs *xs,
                           const monte_mc_spectrum *spec, monte_mc_scene **out);
void monte_gpu_scene_destroy(monte_mc_scene *s);
/* re-upload the label volume of an existing scene (same dimensions) from a host buffer, on `stream` */
int monte_gpu_scene_update_labels(monte_mc_scene *s, const uint8_t *labels, void *stream);
/* Run photons n in [n_begin, n_end) of every pixel of views [view_begin, view_end)
 * and add into d_image0/d_image5 (device, [n_views][ny][nx], caller zeroes).
 * d_stats: device uint64[16] accumulators (nullable), see monte_gpu_mc_stats_unpack. */
/* Launches of ONE scene may overlap on different streams (each takes its own work counter out of a ring of 16);
 * more than 16 launches of one scene in flight at once are not supported.                                  */
int monte_gpu_simulate_dev(const monte_mc_scene *s, uint64_t seed,
                           int view_begin, int view_end,
                           uint32_t n_begin, uint32_t n_end, uint32_t photons_per_pixel,
                           int32_t *d_image0, int32_t *d_image5,
                           unsigned long long *d_stats, void *stream);
#define MONTE_MC_STATS_WORDS 16
void monte_gpu_mc_stats_unpack(const unsigned long long *h_words, monte_mc_stats *out);

/* per-history fate records for history-coupled parity tests (debug; small runs).
 * fate[h] = kind | bin<<8 | n_interactions<<28 ; kind: 1 primary, 2 scatter detected,
 * 3 absorbed, 4 escaped undetected, 5 scatter budget exhausted.                        */
int monte_gpu_simulate_fates(const monte_mc_scene *s, uint64_t seed, int view,
                             uint32_t photons_per_pixel, uint32_t *fates /*host*/,
                             float *energies /*host, nullable*/);

/* counts -> line-integral map: c=min(c,per); c=max(c,1); map=-ln(c)+ln(per)
 * (CBCT_real325im.cu:267-285; monte_cpp/map.cpp:19).                                   */
int monte_gpu_counts_to_map(const int32_t *counts, size_t n, int32_t per, float *map);
int monte_gpu_counts_to_map_dev(const int32_t *d_counts, size_t n, int32_t per,
                                float *d_map, void *stream);

/* Deterministic primary projection: map[view][iy][ix] = integral of mu along the ray
 * from the source to the pixel centre at energy keV (the variance-free limit of
 * -ln(image0/per)); exact voxel traversal.                                             */
int monte_gpu_project_primary(const monte_mc_geom *g, const monte_mc_volume *vol,
                              const uint8_t *labels, const monte_mc_xs *xs,
                              double keV, int view_begin, int view_end, float *map);

/* Device-resident form of the same projector: the label copies are built once, the line integrals are written to
 * a DEVICE buffer d_map [n_views][ny][nx] at the absolute view index -- the layout monte_gpu_fdk_filter_dev reads
 * (iu = iy, iv = ix) -- so projection -> FDK needs no host round trip, and a multi-GPU host projects exactly the
 * views it filters (BASELINE config 3: primary-only projection + FDK, sharded by views then z-slabs).
 * Asynchronous on `stream` (a cudaStream_t passed as void*).                                                   */
typedef struct monte_projector monte_projector;   /* opaque */
int  monte_gpu_projector_create(const monte_mc_volume *vol, const uint8_t *labels /*host*/, monte_projector **out);
void monte_gpu_projector_destroy(monte_projector *p);
int  monte_gpu_project_primary_dev(const monte_projector *p, const monte_mc_geom *g, const monte_mc_xs *xs,
                                   double keV, int view_begin, int view_end, float *d_map, void *stream);

/* ---- host helpers shared by the C++ drivers (no GPU) ---------------------- */
/* readcsv role (CBCT_real2.cpp:633-668): 200 rows "coh,compton,photo,total", UTF-8 BOM
 * and CRLF tolerated, row r -> index r+1 (keV); index 0 is a copy of index 1.
 * quirk_bom != 0 reproduces csvarray[0][1]=1.372 for every file (CBCT_real2.cpp:663).  */
int monte_xs_load_csv(const char *path, int material, float density, int quirk_bom,
                      monte_mc_xs *xs);
/* make_fantom.cpp:10-19: n x n uint8 disc, (j-cy)^2+(k-cx)^2 <= r2 -> 1 */
void monte_make_fantom(uint8_t *g, int n, int cy, int cx, int r2);
/* make_image01.cpp:15-23: nx*ny*nz uint8 sphere */
void monte_make_sphere(uint8_t *g, int nx, int ny, int nz, int cx, int cy, int cz, int r2);
/* ctnum_to_mu role (ctnum_to_mu.cpp:55 computes mu_H2O and stops): HU -> mu(E) =
 * mu_water(E)*(1+HU/1000) and a label segmentation by HU thresholds.                   */
int monte_ctnum_to_mu(const float *hu, size_t n, const monte_mc_xs *xs, double keV,
                      float hu_air_max, float hu_bone_min, float *mu, uint8_t *labels);
/* The same role done so that the TRANSPORT consumes the result: an N-class segmentation of a CT volume.
 * A class covers HU in [hu_min, hu_min of the next class) (the last one is open above); classes ascend in hu_min and
 * everything below the first class is air.  A class with density <= 0 is air as well (label 0); every other class
 * becomes one material of the transport: tables = the mass-fraction mixture (mixture rule: mu/rho = (1-f) (mu/rho)_a
 * + f (mu/rho)_b, per interaction type) of two base materials, density = the class's own -- so one base table serves
 * several density bins (lung / adipose / soft tissue as water at 0.3 / 0.93 / 1.03 g/cm^3) and bone is water + calcium. */
typedef struct monte_hu_class {
    float   hu_min;
    int32_t material_a;      /* index into the base tables (0-based)                                  */
    int32_t material_b;      /* second component, or < 0 for none                                     */
    float   frac_b;          /* mass fraction of material_b, 0..1                                     */
    float   density;         /* g/cm^3 of the class; <= 0: air                                        */
} monte_hu_class;
/* Default table for base tables {0: water (xcom2.csv), 1: calcium (Ca.csv)}: air < -900 | lung | adipose | soft
 * tissue | muscle | spongy bone | bone | dense bone | cortical bone (ICRU-44-like densities and calcium mass
 * fractions 0.05 .. 0.225).  have_calcium == 0: the bone classes are water at their density.  Writes at most
 * MONTE_MC_MAX_MATERIALS + 1 classes; returns their number.                                                        */
int monte_hu_classes_default(int have_calcium, monte_hu_class *classes);
/* hu[n] -> labels[n] (0 = air, k = the k-th non-air class) and the material tables `out` those labels index
 * (out->n_materials = number of non-air classes <= MONTE_MC_MAX_MATERIALS; form-factor tables of material_a are
 * carried over).  mu (nullable): the linear attenuation the transport will see at keV, total[label-1][keV] * density
 * -- the deterministic projector integrates exactly this.  present (nullable): bit m set iff material m occurs.   */
int monte_ctnum_segment(const float *hu, size_t n, const monte_hu_class *classes, int n_classes,
                        const monte_mc_xs *base, double keV, monte_mc_xs *out, uint8_t *labels, float *mu,
                        uint32_t *present);
/* Analytic stand-in for a measured form factor (the reference ships none): F(x)^2 ~ (1 + x^2/x0^2)^-4, the
 * hydrogen-like 1s charge cloud, x0 in 1/Angstrom (0.30 * Z_eff).  Fills ff_x2 / ff_cum of `material` on a
 * logarithmic grid of MONTE_MC_FF_POINTS points up to x^2 = 270 (200 keV back-scatter) and sets ff_points.  */
int monte_xs_formfactor_hydrogenic(monte_mc_xs *xs, int material, double x0);
/* Clearance grid of MONTE_MC_TRACK_CLEARANCE (host only; the library builds it itself at scene upload, the call is
 * exported for hosts that want to inspect it and for the CPU oracle): dims[] = ceil(n / 2^cell_log2) per axis;
 * grid[(cz*dims[1] + cy)*dims[0] + cx] = floor(2 d) clipped to 127, d = smallest distance in cell sides between the
 * cell's box and the box of any cell holding a voxel of `heavy_material` (index into monte_mc_xs, labels above
 * n_materials clamp to the last material as in the transport).  A flight of at most grid * 2^cell_log2 * pitch / 2
 * from anywhere in the cell cannot reach that material.                                                           */
int monte_mc_clearance_dims(const monte_mc_volume *vol, int cell_log2, int32_t dims[3]);
int monte_mc_clearance_grid(const monte_mc_volume *vol, const uint8_t *labels, int n_materials, int heavy_material,
                            int cell_log2, uint8_t *grid);
/* eight grids [octant][cz][cy][cx], octant = (dx > 0) | (dy > 0) << 1 | (dz > 0) << 2 of the flight direction: the same
 * bound restricted to the cells a ray with that direction can still reach (MONTE_MC_TRACK_DIRECTIONAL)            */
int monte_mc_clearance_grid_octants(const monte_mc_volume *vol, const uint8_t *labels, int n_materials,
                                    int heavy_material, int cell_log2, uint8_t *grid8);
/* the material the clearance grid is built for: argmax of total*density at 60 keV; -1 with fewer than 2 materials */
int monte_xs_heavy_material(const monte_mc_xs *xs);
/* What MONTE_MC_TRACK_AUTO resolves to (host only, deterministic in its inputs, so every rank of a sharded run
 * decides alike): the spectrum-weighted mean of mu_max(E) / mu_light(E) -- the factor by which the heavy material
 * inflates the number of tentative collisions in the bulk -- above 3 selects MONTE_MC_TRACK_DIRECTIONAL
 * (*cell_log2 = 2), else MONTE_MC_TRACK_GLOBAL.  Measured on a B200, water + calcium, ms per 1e8 histories (round 2,
 * profiles/r02_mc_tracking_modes.jsonl): 140 keV (ratio 1.8) 8.13 global vs 9.37 directional; 60 keV (5.0) 11.1 vs 8.61;
 * 120 kVp (6.1) 20.9 vs 8.64 (plain CLEARANCE: 9.95).  ratio (nullable) receives the mean.                          */
int monte_mc_resolve_tracking(const monte_mc_xs *xs, const monte_mc_spectrum *spec, int32_t *cell_log2, double *ratio);
/* per-keV Woodcock majorant (1/cm) over the materials that occur in `labels` (NULL: all materials):
 * the max of CBCT_real325im.cu:866 restricted to what the volume contains; mu_max[201].               */
int monte_xs_majorant(const monte_mc_xs *xs, const uint8_t *labels, size_t n, float *mu_max);

#ifdef __cplusplus
}
#endif
#endif /* MONTE_GPU_H */
