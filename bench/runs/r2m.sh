#!/bin/bash
# round 2, run m: 3- and 4-CTA/SM shapes of the production kernel (variants 18-29)
mkdir -p gpurun_out
B2S_LIB=cub_b200/libb2s_tune.so timeout 600 python bench/tune.py --log2n 28 --cases k4v4,k4v0 --variants 0,18,19,20,21,22,23,24,25,26,27,28,29 --out gpurun_out/tune_r2m.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d['case'], d.get('impl'), d.get('variant'), d.get('nt'), d.get('ipt'), d.get('minb'), d.get('flow'), round(d['best_ms'], 3), round(d['gkeys_s'], 2), d.get('bit_exact_vs_ref'))
"
B2S_LIB=cub_b200/libb2s_tune.so timeout 300 python bench/tune.py --log2n 27 --cases k8v4,k8v0 --variants 0,18,21,25,26,27 --out gpurun_out/tune_r2m.jsonl 2>&1 | tail -12 | cut -c1-400
