#!/bin/bash
# round 2, call 18 (1 GPU): thin-slab z-block streams (timing of the 8- and 4-device slabs alone), the tightened MC
# parity bars, the z-stream bit-equality test, MC timing
set -u
mkdir -p gpurun_out
timeout 300 python scripts/fdk_perf.py --iters 2 --slabs 8 > gpurun_out/r02c18_slabs8.log 2>&1
timeout 300 python scripts/fdk_perf.py --iters 2 --slabs 4 > gpurun_out/r02c18_slabs4.log 2>&1
timeout 600 python -m pytest tests/test_fdk_gpu.py tests/test_mc_gpu.py tests/test_tracking_gpu.py tests/test_rayleigh_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/r02c18_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c18_tests.log
tail -n 3 gpurun_out/r02c18_slabs8.log gpurun_out/r02c18_slabs4.log | cut -c1-600; tail -n 4 gpurun_out/r02c18_tests.log
