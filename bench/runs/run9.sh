#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err
echo "bench exit $?"; python -c "
import json; r=json.load(open('gpurun_out/bench_r1h.json')); print(r['value'], r['ms_per_step'], r['roofline']['frac'], r['roofline']['avg_launch_ms'], r['roofline']['histogram_ms'], r['config']['parity'], r['e2e']['value'], r['reference_gpu']['value'])"
timeout 900 python bench/configs.py --out gpurun_out/configs_r1h.jsonl > gpurun_out/configs_h.log 2>&1; echo "configs exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/configs_r1h.jsonl'):
    r=json.loads(l)
    if r['impl']=='b2s': print(r['config'], round(r['ms'],3), round(r['gkeys_s'],2), round(r['hbm_roofline_frac'],3), r['bit_exact_vs_ref'])
PY
