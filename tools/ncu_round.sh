#!/bin/bash
# Round evidence: launch list of one full-size train step + render, and full captures of the dominant kernels.
# Usage (on the GPU box, via gpurun): bash tools/ncu_round.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python tools/profile_step.py > gpurun_out/ncu_${TAG}_a.log 2>&1
# (two captures: the UNet's 40-odd convolution launches would otherwise use up the launch budget before the first weight gradient)
HW=400 WARM=1 RENDER=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"(^|:)(stack_kernel|wgrad_kernel|select_grid_kernel)" -c 26 -o gpurun_out/prof_main_${TAG} python tools/profile_step.py > gpurun_out/ncu_${TAG}_b.log 2>&1
HW=400 WARM=1 RENDER=0 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"(^|:)(conv_kernel|conv_wgrad_kernel)" -c 14 -o gpurun_out/prof_conv_${TAG} python tools/profile_step.py > gpurun_out/ncu_${TAG}_b2.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    -k regex:"stack_kernel|wgrad_kernel|linear_kernel|conv_kernel|conv_wgrad" --csv --log-file gpurun_out/traffic_${TAG}.csv env RENDER=0 python tools/profile_step.py > gpurun_out/ncu_${TAG}_c.log 2>&1
tail -n 1 gpurun_out/ncu_${TAG}_a.log gpurun_out/ncu_${TAG}_b.log gpurun_out/ncu_${TAG}_c.log || true
