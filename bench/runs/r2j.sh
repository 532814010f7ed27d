#!/bin/bash
# round 2, run j: evidence for profiles/: launch list of the bench command, ncu --set full of the production kernels, configs table
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:digit_pass -s 5 -c 1 -o gpurun_out/prof_onesweep_$TAG -f python bench/profile_target.py --reps 2 > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:digit_pass -s 5 -c 1 -o gpurun_out/prof_onesweep_keys_$TAG -f python bench/profile_target.py --reps 2 --case k4v0 >> gpurun_out/ncu_full_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram -s 1 -c 1 -o gpurun_out/prof_hist_$TAG -f python bench/profile_target.py --reps 2 >> gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"
timeout 900 python bench/configs.py --out gpurun_out/configs_$TAG.jsonl > gpurun_out/configs_$TAG.log 2>&1; echo "configs exit $?"; tail -12 gpurun_out/configs_$TAG.log
