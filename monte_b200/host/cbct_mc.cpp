// cbct_mc — the role of the main() of monte_cu/CBCT_real325im.cu (:83-297): read the label volume and
// the cross-section tables, run the photon transport, write the count images and the -log maps in
// the reference's headerless layouts (proj*_0 / proj*_5 int32, map*_0 / map*_5 float32, [view][y][x]).
//   cbct_mc labels.raw N pitch_cm xcom2.csv Ca.csv [det=325] [pixel=0.1] [views=360] [per=10000] [seed=0] [tag=out] [rayleigh=0] [clearance=0] [--gpus G] [--hu] [--kev E]
// --hu: the first file is a CT volume in Hounsfield units (float32) instead of labels: it is segmented into the classes
// of monte_hu_classes_default (density bins of water, water + calcium for bone; the role ctnum_to_mu.cpp:55-82 hints at)
// and tracked with the majorant of the classes that occur (monte_mc_volume.majorant_mode = MONTE_MC_MAJORANT_PRESENT).
// --kev E: source energy (default 140, CBCT_real325im.cu as shipped).
// --ring R: the ring detector of the 2-D programs (monte_cpp/circle3_2.cpp) instead of the flat panel: source at the origin,
// `det` angular bins around the z axis at radius R cm, one axial bin of height `pixel` (images are [views][det][1]).
// --gpus G (1..8): the photons of every pixel are split over G devices inside libmonte_gpu and the tallies summed on
// device 0 (monte_gpu_init(G, NULL)); the output files are byte-identical to --gpus 1.
// rayleigh=1 (not in the reference): coherent events are deflected by the analytic form factor of
// monte_xs_formfactor_hydrogenic (x0 = 1.0 for water, 2.2 for calcium) instead of flying straight on.
// clearance=n > 0 (not in the reference): two-level Woodcock majorant with clearance cells of 2^n voxels
// (monte_mc_volume.tracking_mode = MONTE_MC_TRACK_CLEARANCE): same physics, faster when calcium sets the majorant.
// All compute is in libmonte_gpu (no CPU fallback: the call fails without a B200).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "monte_gpu.h"

static int fail() { fprintf(stderr, "cbct_mc: %s\n", monte_gpu_last_error()); return 1; }
static void write_raw(const std::string &fn, const void *p, size_t bytes) {
    FILE *f = fopen(fn.c_str(), "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "failed to write %s\n", fn.c_str()); exit(1); }
    fclose(f);
}

int main(int argc, char **argv) {
    int gpus = 1;
    bool hu_input = false;
    double kev = 140.0, ring_r = 0.0;
    for (int i = 1; i < argc;) {
        if (!strcmp(argv[i], "--hu")) { hu_input = true; for (int j = i; j + 1 < argc; j++) argv[j] = argv[j + 1]; argc -= 1; continue; }
        const bool gp = !strcmp(argv[i], "--gpus"), ke = !strcmp(argv[i], "--kev"), rg = !strcmp(argv[i], "--ring");
        if ((gp || ke || rg) && i + 1 < argc) {
            if (gp) gpus = atoi(argv[i + 1]); else if (ke) kev = atof(argv[i + 1]); else ring_r = atof(argv[i + 1]);
            for (int j = i; j + 2 < argc; j++) argv[j] = argv[j + 2];
            argc -= 2;
            continue;
        }
        i++;
    }
    if (argc < 6) { fprintf(stderr, "usage: cbct_mc labels.raw N pitch xcom2.csv Ca.csv [det] [pixel] [views] [per] [seed] [tag] [rayleigh] [clearance] [--gpus G]\n"); return 2; }
    const int n = atoi(argv[2]);
    const double pitch = atof(argv[3]);
    const int det = argc > 6 ? atoi(argv[6]) : 325;
    const double pixel = argc > 7 ? atof(argv[7]) : 0.1;
    const int views = argc > 8 ? atoi(argv[8]) : 360;
    const uint32_t per = argc > 9 ? (uint32_t)atol(argv[9]) : 10000;     // num_photon, CBCT_real325im.cu:6
    const uint64_t seed = argc > 10 ? strtoull(argv[10], nullptr, 10) : 0;
    const std::string tag = argc > 11 ? argv[11] : "out";
    const bool rayleigh = argc > 12 && atoi(argv[12]) != 0;
    const int clearance = argc > 13 ? atoi(argv[13]) : 0;
    std::vector<uint8_t> lab((size_t)n * n * n);
    std::vector<float> hu(hu_input ? lab.size() : 0);
    FILE *f = fopen(argv[1], "rb");
    if (!f || (hu_input ? fread(hu.data(), 4, hu.size(), f) != hu.size() : fread(lab.data(), 1, lab.size(), f) != lab.size())) { fprintf(stderr, "failed to read %s\n", argv[1]); return 1; }
    fclose(f);
    std::unique_ptr<monte_mc_xs> xs(new monte_mc_xs());
    if (monte_xs_load_csv(argv[4], 0, 1.0f, 1, xs.get()) || monte_xs_load_csv(argv[5], 1, 1.55f, 1, xs.get())) return fail();
    if (hu_input) {                                             // CT numbers -> labels + the tables they index
        monte_hu_class cls[MONTE_MC_MAX_MATERIALS + 1];
        const int nc = monte_hu_classes_default(1, cls);
        std::unique_ptr<monte_mc_xs> seg(new monte_mc_xs());
        uint32_t present = 0;
        if (nc < 0 || monte_ctnum_segment(hu.data(), hu.size(), cls, nc, xs.get(), kev, seg.get(), lab.data(), nullptr, &present)) return fail();
        printf("segmented into %d classes, present mask 0x%x\n", seg->n_materials, present);
        xs.swap(seg);
    }
    monte_mc_geom g = {};
    g.n_views = views; g.angle0_deg = 0; g.angle_step_deg = 360.0 / views;
    g.ny = g.nx = det; g.pixel = pixel; g.half = 0.5 * det * pixel; g.dso = 160; g.dod = 60;   // :459
    g.source_mode = MONTE_MC_SOURCE_PENCIL; g.max_scatter = 5;                                    // :7
    const int det_x = ring_r > 0 ? 1 : det;                                                       // ring: one axial bin
    if (ring_r > 0) { g.detector_shape = MONTE_MC_DETECTOR_RING; g.ring_radius = ring_r; g.nx = 1; g.half = 0.5 * pixel; }
    if (rayleigh) {
        for (int m = 0; m < xs->n_materials; m++)               // (segmented volumes: water-like classes 1.0, calcium-loaded 1.3)
            if (monte_xs_formfactor_hydrogenic(xs.get(), m, hu_input ? (m < 4 ? 1.0 : 1.3) : (m == 0 ? 1.0 : 2.2))) return fail();
        g.coherent_mode = MONTE_MC_COHERENT_FORMFACTOR;
    }
    monte_mc_volume v = {};
    v.nx = v.ny = v.nz = n; v.pitch = pitch;
    for (int a = 0; a < 3; a++) { v.origin[a] = -0.5 * n * pitch; v.clip_lo[a] = v.origin[a]; v.clip_hi[a] = -v.origin[a]; }
    if (clearance > 0) { v.tracking_mode = MONTE_MC_TRACK_CLEARANCE; v.clearance_cell_log2 = clearance; }
    if (hu_input) v.majorant_mode = MONTE_MC_MAJORANT_PRESENT;
    monte_mc_spectrum sp = {0, 0.5, kev, nullptr};                                                // as shipped: 140 keV
    if (monte_gpu_init(gpus, nullptr)) return fail();
    const size_t n_img = (size_t)views * det * det_x;
    std::vector<int32_t> im0(n_img), im5(n_img);
    std::vector<float> map0(n_img), map5(n_img);
    monte_mc_stats st;
    // launch .. D2H .. clamp + -log maps (CBCT_real325im.cu:232-288) in one call: the maps come out of the pass that
    // sums the per-device tallies
    if (monte_gpu_simulate_maps(&g, &v, lab.data(), xs.get(), &sp, per, 0, per, seed, 0, views, im0.data(), im5.data(),
                                map0.data(), map5.data(), &st)) return fail();
    printf("count = %llu primaries + %llu scattered / %llu histories, %.1f ms on %d GPU%s (%.3g histories/s)\n",
           (unsigned long long)st.primaries, (unsigned long long)st.scatter_detected, (unsigned long long)st.histories,
           st.ms_kernel, gpus, gpus > 1 ? "s" : "", st.histories / (st.ms_kernel * 1e-3));
    write_raw("proj_" + tag + "0.raw", im0.data(), n_img * 4);
    write_raw("proj_" + tag + "5.raw", im5.data(), n_img * 4);
    write_raw("map_" + tag + "0.raw", map0.data(), n_img * 4);
    write_raw("map_" + tag + "5.raw", map5.data(), n_img * 4);
    monte_gpu_shutdown();
    return 0;
}
