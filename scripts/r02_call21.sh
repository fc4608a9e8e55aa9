#!/bin/bash
# round 2, call 21 (1 GPU): ncu launch list of the bench command + ncu --set full of the shipped MC kernel and of the
# backprojector (one view-chunk launch)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02c21_launches.csv python bench.py --steps 4 --warmup 3 --skip-cpu --fdk-steps 1 --skip-c4 --skip-c5 --skip-c1 > gpurun_out/r02c21_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_transport -s 1 -c 1 -o gpurun_out/r02c21_mc python scripts/mc_perf.py --per 947 --views 1 --iters 1 > gpurun_out/r02c21_ncu_mc.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fdk_backproject -s 4 -c 1 -o gpurun_out/r02c21_fdk python scripts/fdk_perf.py --iters 1 > gpurun_out/r02c21_ncu_fdk.log 2>&1
ls -la gpurun_out/r02c21_*; tail -n 2 gpurun_out/r02c21_ncu_mc.log gpurun_out/r02c21_ncu_fdk.log
