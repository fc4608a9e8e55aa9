#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 60 python scripts/mc_perf.py --iters 4 2>&1 | tail -n 1
echo "== poly"; timeout 60 python scripts/mc_perf.py --iters 3 --poly 2>&1 | tail -n 1
echo "== cone"; timeout 60 python scripts/mc_perf.py --iters 3 --cone 2>&1 | tail -n 1
echo "== tests"; timeout 300 python -m pytest tests/test_mc_gpu.py tests/test_tracking_gpu.py tests/test_rayleigh_gpu.py -m gpu -x -q 2>&1 | tail -n 2
