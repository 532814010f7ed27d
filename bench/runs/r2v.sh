#!/bin/bash
# round 2, run v: float shapes at the integer items/thread -- float parity tests + configs table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench/configs.py --out gpurun_out/configs_${1:-r2v}.jsonl 2>&1 | grep '"b2s"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'][:60], round(d['ms'],3), round(d['gkeys_s'],2), round(d['hbm_roofline_frac'],3), d['bit_exact_vs_ref'])"
