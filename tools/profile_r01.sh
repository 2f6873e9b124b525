#!/bin/bash
# Round-1 profiling pass (run under gpurun, ONE GPU): launch list of a short bench + full capture of the SpMV kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r01_clocks.csv &
SMI=$!
# 1. launch list (cold-cache, serialised: compare SHARES): small mesh so the run is short
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r01_launches_n128.csv \
    python bench.py --mesh-n 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r01_launches_bench.log 2>&1
# 2. full capture of the in-solve SpMV (DOT_YX) and the update / direction kernels at the bench size
ncu --set full --clock-control none --import-source on -k regex:"k_spmv_s3_rt|k_cg_update|k_cg_dir" -s 30 -c 6 -o gpurun_out/r01_prof_solve_n256 \
    python bench.py --mesh-n 256 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r01_prof_solve.log 2>&1
kill $SMI
tail -2 gpurun_out/r01_launches_bench.log
