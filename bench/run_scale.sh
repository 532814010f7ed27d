#!/bin/bash
# usage: bash bench/run_scale.sh <ngpus> [log2n for config5]
N=${1:-4}; L=${2:-30}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_mgpu_${N}_peer.json 2> gpurun_out/bench_mgpu_${N}_peer.err
echo "bench exit $?"; cat gpurun_out/bench_mgpu_${N}_peer.json; grep -v "Warning\|^frame\|OMP_NUM\|\*\*\*" gpurun_out/bench_mgpu_${N}_peer.err | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench/mgpu_config5.py --log2n $L --reps 2 > gpurun_out/config5_${N}.jsonl 2> gpurun_out/config5_${N}.err
echo "config5 exit $?"; cat gpurun_out/config5_${N}.jsonl; grep -v "Warning\|^frame\|OMP_NUM\|\*\*\*" gpurun_out/config5_${N}.err | tail -8
