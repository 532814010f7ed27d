#!/bin/bash
# round 2, run e: full GPU test suite on the new production kernel + bench + claim-mode timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
echo "bench exit $?"; python -c "
import json; r=json.load(open('gpurun_out/bench_r2e.json')); print(r['value'], r['ms_per_step'], r['roofline']['frac'], r['roofline'].get('avg_launch_ms'), r['config'].get('parity'), r['e2e']['value'], r.get('reference_gpu',{}).get('value'))"
B2S_TILE_CLAIM=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2e_claim.json 2> gpurun_out/bench_r2e_claim.err
python -c "
import json; r=json.load(open('gpurun_out/bench_r2e_claim.json')); print('claim', r['value'], r['ms_per_step'], r['roofline']['frac'])"
