#!/bin/bash
# Evidence run for the production build: launch list of the bench command, full ncu captures of the two kernels,
# bench line, smoke.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 1 -o gpurun_out/prof_onesweep_r1d -f python bench/profile_target.py --reps 2 > gpurun_out/ncu_full_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram -s 1 -c 1 -o gpurun_out/prof_hist_r1d -f python bench/profile_target.py --reps 2 >> gpurun_out/ncu_full_d.log 2>&1
tail -2 gpurun_out/ncu_full_d.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
echo "bench exit $?"; cat gpurun_out/bench_r1d.json; tail -3 gpurun_out/bench_r1d.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_r1d.json 2> gpurun_out/bench_ref_r1d.err
echo "ref exit $?"; cat gpurun_out/bench_ref_r1d.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
