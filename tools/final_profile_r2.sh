#!/usr/bin/env bash
# Round-2 measurement pass on one B200 (run under gpurun; everything lands in gpurun_out/):
#   1. the bench line with the per-kernel in-graph table (globaltimer stamps)            -> bench_r2.json, per_kernel_ingraph.csv
#   2. ncu launch list of one eager step (cold caches, serialised)                         -> launches_r2.csv
#   3. ncu --set full captures of the dominant kernels of that step (3 launches each)      -> ncu_*_r2.ncu-rep
# Every ncu command runs only after the same command has exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --ops-out gpurun_out/per_kernel_ingraph.csv > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err || exit 1
python tools/profile_step.py > gpurun_out/profile_step_plain.log 2>&1 || { echo "profile_step failed"; exit 1; }
# the profiled step is the LAST 83 launches of the process (warm-up step first)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
    python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
# full captures: kf convs (skip the warm-up step's 39 kf launches), the general conv kernel, the apply kernel, attention
ncu --set full --clock-control none --import-source on -k regex:conv_kf_kernel --launch-skip 39 -c 39 -f \
    -o gpurun_out/ncu_kf_r2 python tools/profile_step.py > gpurun_out/ncu_full_kf.log 2>&1
ncu --set full --clock-control none -k regex:"conv_tc_kernel|gn_apply_kernel|attention_tc_kernel|gn_stats_kernel" --launch-skip 44 -c 44 -f \
    -o gpurun_out/ncu_other_r2 python tools/profile_step.py > gpurun_out/ncu_full_other.log 2>&1
tail -c 400 gpurun_out/bench_r2.json
