#!/bin/bash
# evidence for profiles/ with the summaries extracted ON the box (three --import-source reports exceed the
# 64 MiB that gpurun copies back): launch list of the bench command, ncu --set full of the production kernels, bench lines
TAG=${1:-r2l}
mkdir -p gpurun_out
N32=$((1 << 23))
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_bench_$TAG.log 2>&1
echo "launch list exit $?"
prof() {  # name, kernel regex, skip, case
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 -f python bench/profile_target.py --reps 2 --case $4 >> gpurun_out/ncu_full_$TAG.log 2>&1
  python bench/ncu_summary.py /tmp/$1.ncu-rep $N32 > gpurun_out/$1.ncu.txt 2>&1
  python bench/ncu_by_line.py /tmp/$1.ncu-rep $N32 > gpurun_out/$1.by_line.txt 2>&1
  rm -f /tmp/$1.ncu-rep
}
prof prof_pass_pairs_$TAG digit_pass 5 k4v4
prof prof_pass_keys_$TAG digit_pass 5 k4v0
prof prof_hist_$TAG histogram 1 k4v4
tail -2 gpurun_out/ncu_full_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"
