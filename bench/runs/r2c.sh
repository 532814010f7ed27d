#!/bin/bash
# round 2, run c: ncu --set full of the pair-scatter variant 54 (448x24x2) and of production (variant 0), tuning build
mkdir -p gpurun_out
export B2S_LIB=cub_b200/libb2s_tune.so
for v in 54 0; do
B2S_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 1 -o gpurun_out/prof_r2c_v$v -f python bench/profile_target.py --reps 2 > gpurun_out/ncu_r2c_v$v.log 2>&1
tail -1 gpurun_out/ncu_r2c_v$v.log
done
