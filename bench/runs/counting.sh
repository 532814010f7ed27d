#!/bin/bash
# counting-sort path: parity tests, the A/B table (counting vs digit passes vs reference CUB) with per-launch times, and
# ncu --set full of the joint histogram kernel (summaries extracted on the box)
mkdir -p gpurun_out
tag=${1:-r4a}
keys=${2:-5,4,2,3,0,1}
timeout 900 python -m pytest tests/test_counting_sort_gpu.py -m gpu -x -q 2>&1 | tail -15
rm -f gpurun_out/counting_${tag}.jsonl
fmt='
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d["tag"], d["key"], d["log2n"], d["dist"], "cnt", d["counting_ms"], "dig", d["digit_passes_ms"], "ref", d["ref_cub_ms"], d["bit_exact"], d.get("launch_ms"))'
timeout 900 python bench/counting.py --steps --keys $keys --tag $tag --out gpurun_out/counting_${tag}.jsonl 2>&1 | python -c "$fmt"
if [ -n "$3" ]; then
  N32=$((1 << 24))   # rows of 32 keys at 2^29
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -o /tmp/nh_$tag -f python bench/profile_target.py --reps 3 --case bf16desc --log2n 29 > gpurun_out/ncu_full_nh_$tag.log 2>&1
  python bench/ncu_summary.py /tmp/nh_$tag.ncu-rep $N32 > gpurun_out/prof_$3_$tag.ncu.txt 2>&1
  python bench/ncu_by_line.py /tmp/nh_$tag.ncu-rep $N32 > gpurun_out/prof_$3_$tag.by_line.txt 2>&1
  tail -1 gpurun_out/ncu_full_nh_$tag.log
fi
