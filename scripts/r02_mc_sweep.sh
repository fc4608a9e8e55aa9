#!/bin/bash
# A/B of the MC kernel at C2 (one view, kernel ms): vote thresholds, K variants; then the polyenergetic case
mkdir -p gpurun_out
for t in 16 8 12 20 24; do
  echo "== MONTE_MC_SECOND=$t" ; MONTE_MC_SECOND=$t timeout 60 python scripts/mc_perf.py --iters 4 2>&1 | tail -n 1
done
for k in 34 36 45 46; do
  echo "== MONTE_MC_KERNEL=$k" ; MONTE_MC_KERNEL=$k timeout 60 python scripts/mc_perf.py --iters 3 2>&1 | tail -n 1
done
echo "== poly"; timeout 60 python scripts/mc_perf.py --iters 3 --poly 2>&1 | tail -n 1
echo "== tests"; timeout 300 python -m pytest tests/test_mc_gpu.py -m gpu -x -q 2>&1 | tail -n 2
