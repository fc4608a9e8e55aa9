#!/bin/bash
# round 2, 2-GPU call: full GPU suite (incl. the HU front end, the chunk-interleaved multi-device FDK, the scattered label
# upload), the fate-flip census, bench.py at N = 2 (in-library e2e)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c17_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02c17_gpu_suite.log
timeout 300 python scripts/mc_flip_census.py > gpurun_out/r02c17_flip_census.jsonl 2> gpurun_out/r02c17_flip_census.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c17_bench_n2.json 2> gpurun_out/r02c17_bench_n2.err
tail -n 4 gpurun_out/r02c17_gpu_suite.log; tail -n 2 gpurun_out/r02c17_flip_census.jsonl | cut -c1-300; tail -c 300 gpurun_out/r02c17_bench_n2.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r02c17_bench_n2.json").read().strip().splitlines()[-1])
print("MC", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["cached_labels"]["ms_per_step"])
f=d["fdk"]; print("FDK", f["value"], f["ms_per_step"], f["breakdown_ms"], "e2e", f["e2e"]["value"], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"])
print(d["parity"])
P
