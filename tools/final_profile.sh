#!/usr/bin/env bash
# Round-end measurement pass on one B200: tests, bench line, in-graph class costs, ncu launch list of one step,
# one full ncu capture of the level-0 conv2 (+ fused input GroupNorm: the 3rd kf launch of the profiled step, after 39 of the warm-up step).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
bash tools/gpu_check.sh
timeout 600 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
timeout 900 python tools/graph_cost.py stem final,ddim \
  down.0.rb1.conv1,down.0.rb2.conv1,up.0.rb2.conv1 down.0.rb1.conv2,down.0.rb2.conv2,up.0.rb1.conv2,up.0.rb2.conv2 up.0.rb1.conv1 \
  down.1.rb1.conv2,down.1.rb2.conv,up.1.rb1.conv2,up.1.rb2.conv down.1.rb1.conv1 up.1.rb1.conv1 \
  down.2.rb1.conv2,down.2.rb2.conv,up.2.rb1.conv2,up.2.rb2.conv down.2.rb1.conv1,up.2.rb1.conv1 \
  down.3.rb1.conv2,down.3.rb2.conv,up.3.rb1.conv2,up.3.rb2.conv down.3.rb1.conv1,up.3.rb1.conv1 \
  mid.rb1.conv,mid.rb2.conv qkv,proj .down up.0.up up.1.up up.2.up,up.3.up .res sdpa apply > gpurun_out/graph_cost_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
  python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_kf_kernel --launch-skip 41 -c 1 -f \
  -o gpurun_out/kf_conv2_gn_in python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -c 600 gpurun_out/bench_final.log; cat gpurun_out/graph_cost_final.log
