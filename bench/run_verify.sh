#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 1 -o gpurun_out/prof_onesweep_r1k -f python bench/profile_target.py --reps 2 > gpurun_out/ncu_full_k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 1 -o gpurun_out/prof_onesweep_keys_r1k -f python bench/profile_target.py --reps 2 --case k4v0 >> gpurun_out/ncu_full_k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram -s 1 -c 1 -o gpurun_out/prof_hist_r1k -f python bench/profile_target.py --reps 2 >> gpurun_out/ncu_full_k.log 2>&1
tail -2 gpurun_out/ncu_full_k.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err
echo "bench exit $?"; python -c "
import json; r=json.load(open('gpurun_out/bench_r1k.json')); print(r['value'], r['ms_per_step'], r['roofline']['frac'], r['roofline']['avg_launch_ms'], r['roofline']['histogram_ms'], r['config']['parity'], r['e2e']['value'], r['reference_gpu']['value'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r1k.json 2> gpurun_out/bench_ref_r1k.err; echo "ref exit $?"
timeout 900 python bench/configs.py --out gpurun_out/configs_r1k.jsonl > gpurun_out/configs_k.log 2>&1; echo "configs exit $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
