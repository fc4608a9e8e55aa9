#!/bin/bash
for k in 35 34 44 45 46; do
  echo "== MONTE_MC_KERNEL=$k" ; MONTE_MC_KERNEL=$k timeout 60 python scripts/mc_perf.py --iters 3 2>&1 | tail -n 1 | cut -c1-140
done
for t in 12 14 18; do
  echo "== MONTE_MC_SECOND=$t" ; MONTE_MC_SECOND=$t timeout 60 python scripts/mc_perf.py --iters 3 2>&1 | tail -n 1 | cut -c1-140
done
