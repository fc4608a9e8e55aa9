#!/bin/bash
# round 2, call 41 (8 GPUs): the full bench line at N = 8 with the final build, then N = 4 reduced
set -u
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --skip-cpu > gpurun_out/r02c41_bench_n8.json 2> gpurun_out/r02c41_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 4 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c41_bench_n4.json 2> gpurun_out/r02c41_bench_n4.err
tail -c 200 gpurun_out/r02c41_bench_n8.err
python - <<'P'
import json
for n in ("8", "4"):
    try:
        d=json.loads(open("gpurun_out/r02c41_bench_n%s.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    f=d["fdk"]
    print("N=%s MC %.4g (%.3f ms) e2e %.4g (%.3f ms) | FDK %.0f GUPS (%.2f ms) %s | e2e %.2f ms %s | parity %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
          f["value"], f["ms_per_step"], f["breakdown_ms"], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"], d["parity"].get("ok")))
    if d.get("fdk_c5"): print("   C5", [(p["volume"], round(p["gups"]), round(p["ms_per_reconstruction"],2)) for p in d["fdk_c5"]["points"]])
    if d.get("mc_c4"): print("   C4", d["mc_c4"]["value"], d["mc_c4"]["seconds_for_the_run"])
P
