#!/bin/bash
# round 2, GPU call 3 (2 GPUs): the reference's CUDA kernel beside ours; bench.py at N = 1 and N = 2 (both arms)
set -u
mkdir -p gpurun_out
timeout 700 python scripts/ref_cuda_run.py --out gpurun_out/r02c3_ref_cuda.json > gpurun_out/r02c3_ref_cuda.log 2>&1
timeout 600 python bench.py > gpurun_out/r02c3_bench_n1.json 2> gpurun_out/r02c3_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02c3_bench_n2.json 2> gpurun_out/r02c3_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 --c1-views 36 > gpurun_out/r02c3_bench_n2_ref.json 2> gpurun_out/r02c3_bench_n2_ref.err
tail -c 600 gpurun_out/r02c3_*.err; tail -c 300 gpurun_out/r02c3_*.json
