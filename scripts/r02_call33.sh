#!/bin/bash
# round 2, call 33 (1 GPU): final GPU suite + smoke, compute-sanitizer memcheck over the round's new kernels, bench.py N = 1
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c33_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02c33_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c33_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02c33_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mc_gpu.py tests/test_fdk_gpu.py -m gpu -x -q -k "ring or hu_volume or all_air or update_labels or segment or two_d or chunking or slabs_equal or edge" > gpurun_out/r02c33_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02c33_memcheck.log
timeout 900 python bench.py > gpurun_out/r02c33_bench_n1.json 2> gpurun_out/r02c33_bench_n1.err
tail -n 3 gpurun_out/r02c33_gpu_suite.log; tail -n 2 gpurun_out/r02c33_smoke.log | cut -c1-400; tail -n 6 gpurun_out/r02c33_memcheck.log | cut -c1-300; tail -c 300 gpurun_out/r02c33_bench_n1.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02c33_bench_n1.json").read().strip().splitlines()[-1])
f=d["fdk"]
print("MC %.4g (%.3f ms) e2e %.4g (%.3f ms) roof %.3f util %.3f | FDK %.0f (%.2f ms) e2e %.0f (%.2f ms) | cpu %.3g | parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["issue_util"], f["value"], f["ms_per_step"], f["e2e"]["value"], f["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["parity"].get("ok")))
P
