#!/bin/bash
# round 2, call 19 (1 GPU): the full GPU suite with the ring detector, the 2-D fan-beam case, the reference-CUDA chi-square
# test; MC timing (the default kernel must be untouched by the ring instantiation); bench.py at N = 1
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c19_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02c19_gpu_suite.log
timeout 120 python scripts/mc_perf.py --iters 4 > gpurun_out/r02c19_mc.log 2>&1
timeout 900 python bench.py > gpurun_out/r02c19_bench_n1.json 2> gpurun_out/r02c19_bench_n1.err
tail -n 4 gpurun_out/r02c19_gpu_suite.log; tail -n 1 gpurun_out/r02c19_mc.log | cut -c1-200; tail -c 300 gpurun_out/r02c19_bench_n1.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r02c19_bench_n1.json").read().strip().splitlines()[-1])
print("MC", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"].get("issue_util"))
f=d["fdk"]; print("FDK", f["value"], f["ms_per_step"], "e2e", f["e2e"]["value"], f["e2e"]["ms_per_step"])
print(d["cpu_baseline"]["value"], d["parity"])
P
