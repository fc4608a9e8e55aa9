#!/bin/bash
# First GPU call of round 2 (about 60 s of box time on one B200; everything it runs was verified on the CPU only,
# or changed after its last GPU run, in round 1 — see profiles/README.md, last section):
#   gpurun --timeout 120 -- 'bash scripts/r02_measure.sh'
# Writes gpurun_out/r02_*.log.  Nothing here is a bench value; bench.py is.
set -u
mkdir -p gpurun_out
# 1. the GPU tests written or touched after the last GPU session (tracking modes incl. ADAPTIVE, resident projector)
timeout 60 python -m pytest tests/test_tracking_gpu.py tests/test_rayleigh_gpu.py -m gpu -x -q > gpurun_out/r02_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02_tests.log
# 2. transport modes at C2: reference loop / CLEARANCE cells 4,8,16 / Rayleigh (scripts/mc_modes_perf.py), then ADAPTIVE
#    and DIRECTIONAL (the candidate for the C2 headline: 21 % fewer tentative collisions at 140 keV under emulation)
timeout 60 python scripts/mc_modes_perf.py > gpurun_out/r02_mc_modes.log 2>&1
timeout 60 python - > gpurun_out/r02_mc_adaptive.log 2>&1 <<'PY'
import json, sys
sys.path.insert(0, ".")
from monte_b200 import _abi, api, scenes
api.init(0)
g, vol, lab = scenes.config_c2()
xs = scenes.make_xs()
poly, keep = scenes.kramers_spectrum()
for name, spec in (("mono140", scenes.mono_spectrum(140.0)), ("mono60", scenes.mono_spectrum(60.0)), ("kramers120", poly)):
    for mode, cl in ((_abi.TRACK_GLOBAL, 0), (_abi.TRACK_CLEARANCE, 2), (_abi.TRACK_ADAPTIVE, 2), (_abi.TRACK_ADAPTIVE, 3),
                     (_abi.TRACK_DIRECTIONAL, 2), (_abi.TRACK_DIRECTIONAL, 3)):
        vol.tracking_mode, vol.clearance_cell_log2 = mode, cl
        best = 1e30
        for it in range(3):
            _, _, st = api.simulate(g, vol, lab, xs, spec, 947, seed=it + 1, views=(it, it + 1))
            if it:
                best = min(best, st["ms_kernel"])
        print(json.dumps({"spectrum": name, "tracking_mode": mode, "cell_log2": cl, "ms_kernel": best,
                          "steps_per_hist": st["woodcock_steps"] / st["histories"]}), flush=True)
PY
# 2b. the headline through bench.py with the directional majorant (same physics; compare ms_per_step with the default run)
timeout 60 python bench.py --skip-cpu --skip-fdk --skip-c4 --steps 8 --mc-tracking 4 --mc-cell-log2 3 > gpurun_out/r02_bench_tracking4.log 2>&1
# 3. deterministic projector at C3, host-buffer call: voxel walk vs macro-cells / leaping (one process per setting)
for m in 0 2 3 4; do
  MONTE_PROJ_MACRO=$m timeout 90 python scripts/c3_pipeline.py > gpurun_out/r02_c3_macro$m.log 2>&1
done
# 4. config 3 resident on the device (projection -> filter -> backprojection, no host round trip)
timeout 90 python scripts/c3_pipeline_resident.py > gpurun_out/r02_c3_resident.log 2>&1
MONTE_PROJ_MACRO=3 timeout 90 python scripts/c3_pipeline_resident.py > gpurun_out/r02_c3_resident_macro3.log 2>&1
tail -n 3 gpurun_out/r02_*.log
