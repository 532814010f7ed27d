#!/bin/bash
# new counting-path tests + ncu --set full of the path's kernels (bf16, 2^29 keys, +-0 present), summaries extracted on the box
mkdir -p gpurun_out
tag=${1:-r4g}
timeout 900 python -m pytest tests/test_counting_sort_gpu.py -m gpu -x -q 2>&1 | tail -5
N32=$((1 << 24))
for k in joint_hist16 expand_kernel zero_count16 zero_write16 prefix16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o /tmp/nc_$k -f python bench/profile_target.py --reps 3 --case bf16desc --log2n 29 > gpurun_out/ncu_full_${k}_$tag.log 2>&1
  python bench/ncu_summary.py /tmp/nc_$k.ncu-rep $N32 2>&1 | head -60 > gpurun_out/prof_${k}_$tag.ncu.txt
  rm -f /tmp/nc_$k.ncu-rep
  grep -E "gpu__time_duration|dram__bytes_read.sum|dram__bytes_write.sum|lsu_wavefronts.avg.pct|issue_active" gpurun_out/prof_${k}_$tag.ncu.txt | tr '\n' ' '; echo
done
