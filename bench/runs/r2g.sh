#!/bin/bash
# round 2, run g: 8-GPU scaling of the native multi-GPU sorter (+ config 5 leg), after the faster split_count kernel
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2g_n${N}.json 2> gpurun_out/bench_r2g_n${N}.err
echo "bench exit $?"; wc -l gpurun_out/bench_r2g_n${N}.json; grep -v "Warning\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_r2g_n${N}.err | tail -5
