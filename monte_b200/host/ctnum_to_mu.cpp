// ctnum_to_mu — CT numbers (HU, float32 raw) -> linear attenuation mu(E) and material labels.
// The reference program of this name loads xcom2.csv, computes mu_H2O = csv[3][(int)(E+0.5)]*rho
// (monte_cpp/ctnum_to_mu.cpp:55) and then only crops a volume; the conversion it is named after is
// mu = mu_water(E) * (1 + HU/1000), done here with the same table.
//   ctnum_to_mu hu.raw n_voxels xcom2.csv [keV=140] [mu_out=mu.raw] [label_out=labels.raw] [Ca.csv]
// With Ca.csv the volume is segmented into the classes of monte_hu_classes_default (monte_ctnum_segment: lung / adipose /
// soft tissue / muscle as water at their densities, four bone classes as water + calcium): labels.raw is then what
// cbct_mc transports (cbct_mc --hu does the same in one step) and mu.raw the attenuation the transport sees at keV.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "monte_gpu.h"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ctnum_to_mu hu.raw n_voxels xcom2.csv [keV] [mu.raw] [labels.raw]\n"); return 2; }
    const size_t n = strtoull(argv[2], nullptr, 10);
    const double keV = argc > 4 ? atof(argv[4]) : 140.0;
    std::unique_ptr<monte_mc_xs> xs(new monte_mc_xs());
    if (monte_xs_load_csv(argv[3], 0, 1.0f, 1, xs.get())) { fprintf(stderr, "%s\n", monte_gpu_last_error()); return 1; }
    std::vector<float> hu(n), mu(n);
    std::vector<uint8_t> lab(n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(hu.data(), sizeof(float), n, f) != n) { fprintf(stderr, "failed to read %s\n", argv[1]); return 1; }
    fclose(f);
    if (argc > 7) {
        if (monte_xs_load_csv(argv[7], 1, 1.55f, 1, xs.get())) { fprintf(stderr, "%s\n", monte_gpu_last_error()); return 1; }
        monte_hu_class cls[MONTE_MC_MAX_MATERIALS + 1];
        const int nc = monte_hu_classes_default(1, cls);
        std::unique_ptr<monte_mc_xs> seg(new monte_mc_xs());
        uint32_t present = 0;
        if (nc < 0 || monte_ctnum_segment(hu.data(), n, cls, nc, xs.get(), keV, seg.get(), lab.data(), mu.data(), &present)) { fprintf(stderr, "%s\n", monte_gpu_last_error()); return 1; }
        for (int c = 0; c < nc; c++)
            printf("class %d: HU >= %g, density %.3g g/cm3, calcium mass fraction %.3g, mu(%g keV) = %.5g /cm%s\n", c + 1, cls[c].hu_min, cls[c].density,
                   cls[c].material_b >= 0 ? cls[c].frac_b : 0.f, keV, seg->total[c][(int)(keV + 0.5)] * seg->density[c], (present >> c & 1u) ? "" : "  (absent)");
    } else
    if (monte_ctnum_to_mu(hu.data(), n, xs.get(), keV, -500.f, 700.f, mu.data(), lab.data())) { fprintf(stderr, "%s\n", monte_gpu_last_error()); return 1; }
    printf("mu_H2O(%g keV) = %g /cm\n", keV, xs->total[0][(int)(keV + 0.5)] * xs->density[0]);
    f = fopen(argc > 5 ? argv[5] : "mu.raw", "wb");
    fwrite(mu.data(), sizeof(float), n, f); fclose(f);
    f = fopen(argc > 6 ? argv[6] : "labels.raw", "wb");
    fwrite(lab.data(), 1, n, f); fclose(f);
    return 0;
}
