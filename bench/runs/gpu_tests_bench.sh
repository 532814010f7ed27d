#!/bin/bash
# GPU tests + bench + configs table with the 256-thread pair shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${1:-r2q}.json 2> gpurun_out/bench_${1:-r2q}.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${1:-r2q}.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"])
PY
timeout 900 python bench/configs.py --out gpurun_out/configs_${1:-r2q}.jsonl 2>&1 | grep '"b2s"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'][:60], round(d['ms'],3), round(d['gkeys_s'],2), round(d['hbm_roofline_frac'],3), d['bit_exact_vs_ref'])"
