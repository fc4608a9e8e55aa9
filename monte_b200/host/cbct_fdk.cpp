// cbct_fdk — the role of the main() of recon/bp3d20.cpp / bp3d20_325.cpp / fbp2.cpp: read a float32
// map [views][nu][nv], reconstruct, write xy / zy volumes and the filtered map (bp3d20.cpp:170-190).
//   cbct_fdk bp3d20|bp3d20_325|fbp2 map.raw [tag=out] [--full] [--gpus G]   (--full: whole volume, not s in [125,130))
// --gpus G (1..8): filter sharded by views, backprojection by z-slabs over G devices inside libmonte_gpu
// (monte_gpu_init(G, NULL)); the output files are byte-identical to --gpus 1.  fbp2 (2-D) runs on one device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "monte_gpu.h"

static int fail() { fprintf(stderr, "cbct_fdk: %s\n", monte_gpu_last_error()); return 1; }
static void write_raw(const std::string &fn, const void *p, size_t bytes) {
    FILE *f = fopen(fn.c_str(), "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "failed to write %s\n", fn.c_str()); exit(1); }
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: cbct_fdk bp3d20|bp3d20_325|fbp2 map.raw [tag] [--full]\n"); return 2; }
    const std::string prog = argv[1], tag = argc > 3 && argv[3][0] != '-' ? argv[3] : "out";
    bool full = false;
    int gpus = 1;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--full")) full = true;
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[i + 1]);
    }
    monte_fdk_geom g;
    if (prog == "bp3d20") monte_fdk_geom_bp3d20(&g);
    else if (prog == "bp3d20_325") monte_fdk_geom_bp3d20_325(&g);
    else if (prog == "fbp2") monte_fdk_geom_fbp2(&g);
    else { fprintf(stderr, "unknown program %s\n", prog.c_str()); return 2; }
    if (full) { g.s_begin = 0; g.s_end = g.nx; }
    const size_t n_map = (size_t)g.n_views * g.nu * g.nv, n_vol = (size_t)g.nx * g.ny * g.nz;
    std::vector<float> map(n_map), filt(n_map), xy(n_vol), zy(n_vol);
    FILE *f = fopen(argv[2], "rb");
    if (!f || fread(map.data(), sizeof(float), n_map, f) != n_map) { fprintf(stderr, "failed to read %s\n", argv[2]); return 1; }
    fclose(f);
    if (monte_gpu_init(prog == "fbp2" ? 1 : gpus, nullptr)) return fail();
    monte_fdk_stats st;
    if (prog == "fbp2") {
        if (monte_gpu_fbp2(&g, 1, map.data(), filt.data(), xy.data(), &st)) return fail();
    } else {
        if (monte_gpu_fdk(&g, map.data(), filt.data(), xy.data(), zy.data(), &st)) return fail();
        write_raw("zy_" + tag + ".raw", zy.data(), n_vol * 4);
    }
    printf("filtered\n%.1f ms total on the GPU, %.3g voxel-updates/s\n", st.ms_total,
           st.voxel_updates / ((st.ms_backproject > 0 ? st.ms_backproject : st.ms_total) * 1e-3));
    write_raw("xy_" + tag + ".raw", xy.data(), n_vol * 4);
    write_raw("map_" + tag + ".raw", filt.data(), n_map * 4);
    monte_gpu_shutdown();
    return 0;
}
