#!/bin/bash
# round 2, 8-GPU call: the multi-device C ABI at 8 devices, bench.py at N = 8 (in-library e2e, parity block, C4 at 1e11)
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c16_gpus.log 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02c16_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02c16_multi.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --skip-cpu > gpurun_out/r02c16_bench_n8.json 2> gpurun_out/r02c16_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --skip-cpu --skip-c5 --skip-c4 > gpurun_out/r02c16_bench_n4.json 2> gpurun_out/r02c16_bench_n4.err
tail -n 3 gpurun_out/r02c16_multi.log; tail -c 400 gpurun_out/r02c16_bench_n8.err; tail -c 1500 gpurun_out/r02c16_bench_n8.json
