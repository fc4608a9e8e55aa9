#!/bin/bash
# round 2, GPU call 2 (2 GPUs): the multi-device C ABI on hardware, the full GPU suite, the reference's own CUDA kernel
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c2_gpus.log 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02c2_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r02c2_multi.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02c2_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02c2_gpu_suite.log
timeout 400 python scripts/ref_cuda_run.py --variant im_s --ncu --out gpurun_out/r02c2_ref_cuda_im_s.json > gpurun_out/r02c2_ref_cuda_im_s.log 2>&1
timeout 300 python scripts/ref_cuda_run.py --variant s --out gpurun_out/r02c2_ref_cuda_s.json > gpurun_out/r02c2_ref_cuda_s.log 2>&1
tail -n 3 gpurun_out/r02c2_*.log
