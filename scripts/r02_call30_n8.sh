#!/bin/bash
# round 2, call 30 (8 GPUs): the peer-rows exchange with the pipelined gather at N = 8 and 4; band all_to_all at N = 8 beside it
set -u
mkdir -p gpurun_out
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c30_bench_n$n.json 2> gpurun_out/r02c30_bench_n$n.err
done
MONTE_BENCH_FDK_EXCHANGE=band timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --skip-cpu --skip-c5 --skip-c4 --skip-c1 > gpurun_out/r02c30_bench_n8_band.json 2> gpurun_out/r02c30_bench_n8_band.err
python - <<'P'
import json
for n in ("8", "4", "8_band"):
    try:
        d=json.loads(open("gpurun_out/r02c30_bench_n%s.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    f=d["fdk"]
    print("N=%s MC %.4g e2e %.4g | FDK %.0f GUPS (%.2f ms) %s | e2e %.2f ms %s | parity %s" % (n, d["value"], d["e2e"]["value"],
          f["value"], f["ms_per_step"], f["breakdown_ms"], f["e2e"]["ms_per_step"], f["e2e"]["breakdown_ms"], d["parity"].get("ok")))
P
