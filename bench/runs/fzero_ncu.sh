#!/bin/bash
# ncu --set full of the four digit passes of one f32 keys-only sort with zero recording (first / last: ImageFloatOp, middle: integer
# kernels) and of the two restore kernels; summaries extracted on the box
mkdir -p gpurun_out
tag=${1:-r4p}
N32=$((1 << 23))
timeout 900 ncu --set full --clock-control none -k regex:"digit_pass|fzero" -s 6 -c 6 -o /tmp/fz_$tag -f python bench/profile_target.py --reps 2 --case f32desc --log2n 28 > gpurun_out/ncu_full_fz_$tag.log 2>&1
python bench/ncu_summary.py /tmp/fz_$tag.ncu-rep $N32 2>&1 | grep -v "^   [A-Z0-9.]* .*% of instr\|hot instructions\|^ *[0-9]* *[0-9.]*%  " > gpurun_out/prof_fzero_passes_$tag.ncu.txt
grep -E "Kernel Name|gpu__time_duration|dram__bytes_read.sum|dram__bytes_write.sum|issue_active|lsu_wavefronts.avg.pct|warp-instructions" gpurun_out/prof_fzero_passes_$tag.ncu.txt | cut -c1-200
