#!/bin/bash
# round 2, final 8-GPU call: the multi-device tests at 8 devices, bench.py at N = 8 (full), 4 and 2 (C3 + C2 only)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_dist_gpu.py -m gpu -x -q > gpurun_out/r02c24_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02c24_multi.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --skip-cpu > gpurun_out/r02c24_bench_n8.json 2> gpurun_out/r02c24_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c24_bench_n4.json 2> gpurun_out/r02c24_bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c24_bench_n2.json 2> gpurun_out/r02c24_bench_n2.err
tail -n 3 gpurun_out/r02c24_multi.log
python - <<'P'
import json
for n in (8, 4, 2):
    try:
        d=json.loads(open("gpurun_out/r02c24_bench_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    f=d["fdk"]
    print("N=%d MC %.4g (%.3f ms) e2e %.4g (%.3f ms) | FDK %.0f GUPS (%.2f ms) %s | e2e %.0f (%.2f ms) %s | parity %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
          f["value"], f["ms_per_step"], f["breakdown_ms"], f["e2e"]["value"], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"], d["parity"].get("ok")))
P
