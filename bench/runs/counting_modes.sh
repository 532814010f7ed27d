#!/bin/bash
# A/B of the joint histogram's inner loop on the TUNING library (the product carries the production mode only) (B2S_NH_MODE: bit 0 = branch-free atomics, bit 1 = prefetch, bit 2 = one dummy word per warp)
mkdir -p gpurun_out
tag=${1:-r4d}
modes=${2:-"0 1 2 3"}
rm -f gpurun_out/counting_modes_${tag}.jsonl
fmt='
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d["tag"], d["key"], d["log2n"], d["dist"], "cnt", d["counting_ms"], d["bit_exact"], d.get("launch_ms"))'
for m in $modes; do
  B2S_LIB=$PWD/cub_b200/libb2s_tune.so B2S_NH_MODE=$m timeout 600 python bench/counting.py --steps --keys 5,2 --min-log2 27 --tag mode$m --out gpurun_out/counting_modes_${tag}.jsonl 2>&1 | python -c "$fmt"
done
