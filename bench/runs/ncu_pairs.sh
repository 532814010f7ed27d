#!/bin/bash
# ncu --set full of the 256 x 46 x 2 pair kernel (summaries extracted on the box)
TAG=${1:-r2r}
mkdir -p gpurun_out
N32=$((1 << 23))
timeout 600 ncu --set full --clock-control none --import-source on -k regex:digit_pass -s 5 -c 1 -o /tmp/p_$TAG -f python bench/profile_target.py --reps 2 --case k4v4 > gpurun_out/ncu_full_$TAG.log 2>&1
python bench/ncu_summary.py /tmp/p_$TAG.ncu-rep $N32 > gpurun_out/prof_pass_pairs_$TAG.ncu.txt 2>&1
python bench/ncu_by_line.py /tmp/p_$TAG.ncu-rep $N32 > gpurun_out/prof_pass_pairs_$TAG.by_line.txt 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log
