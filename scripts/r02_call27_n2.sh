#!/bin/bash
# round 2, call 27 (2 GPUs): the CUDA-IPC peer-rows form of the sharded FDK (NCCL worker test), multi-device tests, bench at N = 2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_multi_gpu.py tests/test_fdk_gpu.py -m gpu -x -q > gpurun_out/r02c27_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c27_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --skip-cpu --skip-c4 --skip-c1 > gpurun_out/r02c27_bench_n2.json 2> gpurun_out/r02c27_bench_n2.err
tail -n 5 gpurun_out/r02c27_tests.log; tail -c 400 gpurun_out/r02c27_bench_n2.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02c27_bench_n2.json").read().strip().splitlines()[-1])
f=d["fdk"]
print("MC %.4g e2e %.4g | FDK %.0f GUPS (%.2f ms) %s | %s | e2e %.2f ms %s | c5 %s | parity %s" % (d["value"], d["e2e"]["value"], f["value"], f["ms_per_step"], f["breakdown_ms"], f["config"]["parallelism"][-160:], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"], [(p["volume"], round(p["gups"])) for p in d["fdk_c5"]["points"]], d["parity"]))
P
