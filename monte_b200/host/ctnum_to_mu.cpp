// ctnum_to_mu — CT numbers (HU, float32 raw) -> linear attenuation mu(E) and material labels.
// The reference program of this name loads xcom2.csv, computes mu_H2O = csv[3][(int)(E+0.5)]*rho
// (monte_cpp/ctnum_to_mu.cpp:55) and then only crops a volume; the conversion it is named after is
// mu = mu_water(E) * (1 + HU/1000), done here with the same table.
//   ctnum_to_mu hu.raw n_voxels xcom2.csv [keV=140] [mu_out=mu.raw] [label_out=labels.raw]
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
    if (monte_ctnum_to_mu(hu.data(), n, xs.get(), keV, -500.f, 700.f, mu.data(), lab.data())) { fprintf(stderr, "%s\n", monte_gpu_last_error()); return 1; }
    printf("mu_H2O(%g keV) = %g /cm\n", keV, xs->total[0][(int)(keV + 0.5)] * xs->density[0]);
    f = fopen(argc > 5 ? argv[5] : "mu.raw", "wb");
    fwrite(mu.data(), sizeof(float), n, f); fclose(f);
    f = fopen(argc > 6 ? argv[6] : "labels.raw", "wb");
    fwrite(lab.data(), 1, n, f); fclose(f);
    return 0;
}
