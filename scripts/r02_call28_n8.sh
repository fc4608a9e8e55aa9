#!/bin/bash
# round 2, call 28 (8 GPUs): bench.py at N = 8 (full) and N = 4 with the peer-rows FDK exchange and the chunk-major enqueue
set -u
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --skip-cpu > gpurun_out/r02c28_bench_n8.json 2> gpurun_out/r02c28_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 4 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c28_bench_n4.json 2> gpurun_out/r02c28_bench_n4.err
tail -c 300 gpurun_out/r02c28_bench_n8.err
python - <<'P'
import json
for n in (8, 4):
    try:
        d=json.loads(open("gpurun_out/r02c28_bench_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    f=d["fdk"]
    print("N=%d MC %.4g (%.3f ms) e2e %.4g (%.3f ms) | FDK %.0f GUPS (%.2f ms) %s | e2e %.0f (%.2f ms) %s | parity %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
          f["value"], f["ms_per_step"], f["breakdown_ms"], f["e2e"]["value"], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"], d["parity"].get("ok")))
    if d.get("fdk_c5"): print("   C5", [(p["volume"], round(p["gups"]), round(p["ms_per_reconstruction"],2)) for p in d["fdk_c5"]["points"]])
    if d.get("mc_c4"): print("   C4", d["mc_c4"]["value"], d["mc_c4"]["seconds_for_the_run"])
P
