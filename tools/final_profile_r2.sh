#!/usr/bin/env bash
# Round-2 measurement pass on one B200 (run under gpurun; everything lands in gpurun_out/, keep it under 64 MiB):
#   1. the bench line with the per-kernel in-graph table (globaltimer stamps)            -> bench_r2.json, per_kernel_ingraph.csv
#   2. ncu launch list of one eager step (cold caches, serialised)                         -> launches_r2.csv
#   3. ncu --set full captures of the distinct kernels of that step                        -> ncu_*_r2.ncu-rep
# Every ncu command runs only after the same command has exited 0 without ncu (B200_PROFILING.md).
# Launch order of a step (tools/profile_step.py runs a warm-up step first: 83 launches, 39 of them conv_kf_kernel).
set -u
mkdir -p gpurun_out
if [ "${1:-all}" != "ncu" ]; then
  python bench.py --steps 10 --warmup 3 --ops-out gpurun_out/per_kernel_ingraph.csv > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err || exit 1
fi
python tools/profile_step.py > gpurun_out/profile_step_plain.log 2>&1 || { echo "profile_step failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
    python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
# kf convs: the first 17 of the step (stem, level-0 plain / fused-GroupNorm, stride-2, level-1 +res / fused / plain, BN 48, BN 32)
ncu --set full --clock-control none --import-source on -k regex:conv_kf_kernel --launch-skip 39 -c 17 -f \
    -o gpurun_out/ncu_kf_head_r2 python tools/profile_step.py > gpurun_out/ncu_full_kf1.log 2>&1
# ... and the last 11 (nearest-x2 convs, the up path's +res convs, final conv + DDIM)
ncu --set full --clock-control none --import-source on -k regex:conv_kf_kernel --launch-skip 67 -c 11 -f \
    -o gpurun_out/ncu_kf_tail_r2 python tools/profile_step.py > gpurun_out/ncu_full_kf2.log 2>&1
ncu --set full --clock-control none -k regex:conv_tc_kernel --launch-skip 13 -c 6 -f \
    -o gpurun_out/ncu_tc_r2 python tools/profile_step.py > gpurun_out/ncu_full_tc.log 2>&1
ncu --set full --clock-control none -k regex:"gn_apply_kernel|gn_stats_kernel" --launch-skip 31 -c 14 -f \
    -o gpurun_out/ncu_gn_r2 python tools/profile_step.py > gpurun_out/ncu_full_gn.log 2>&1
ncu --set full --clock-control none -k regex:attention_tc_kernel --launch-skip 1 -c 1 -f \
    -o gpurun_out/ncu_attn_r2 python tools/profile_step.py > gpurun_out/ncu_full_attn.log 2>&1
# summarise on the box and drop the bulky reports (gpurun_out/ must stay under 64 MiB); one small report is kept whole
for r in kf_head kf_tail tc gn attn; do
  python tools/ncu_summary.py gpurun_out/ncu_${r}_r2.ncu-rep gpurun_out/ncu_${r}_r2.csv
done
rm -f gpurun_out/ncu_kf_head_r2.ncu-rep gpurun_out/ncu_kf_tail_r2.ncu-rep gpurun_out/ncu_tc_r2.ncu-rep gpurun_out/ncu_gn_r2.ncu-rep
ls -la gpurun_out | tail -20
tail -c 300 gpurun_out/bench_r2.json
