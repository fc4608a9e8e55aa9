#!/bin/bash
# round 2, call 37 (8 GPUs): bench.py at N = 8 (reduced: C2 + C3) with the final multi-device paths
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c37_bench_n8.json 2> gpurun_out/r02c37_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c37_bench_n4.json 2> gpurun_out/r02c37_bench_n4.err
python - <<'P'
import json
for n in ("8", "4"):
    try:
        d=json.loads(open("gpurun_out/r02c37_bench_n%s.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    f=d["fdk"]
    print("N=%s MC %.4g e2e %.4g (%.3f ms; cached %.3f) | FDK %.0f GUPS (%.2f ms) %s | e2e %.2f ms %s | parity %s" % (n, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["cached_labels"]["ms_per_step"],
          f["value"], f["ms_per_step"], f["breakdown_ms"], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"], d["parity"].get("ok")))
P
