#!/bin/bash
# The multi-GPU checks the driver's single-GPU test box cannot run (NCCL refuses two ranks on one device).  Under gpurun --gpus N:
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/multi_gpu_checks.sh 2'
# 1. the NCCL tests of tests/test_gpu_multi.py (slab broad phase, slab solver with the caller's collective, the library's own NCCL step)
# 2. bench.py at N GPUs under torch.distributed.run: replica headline + the partition block (spheres1m x-slabs, ragdolls5k islands)
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "=== NCCL tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "two_gpu or partitioned" 2>&1 | tail -6
echo "=== bench.py at $N GPUs"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --cpu-steps 1 \
    > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -c 6000 gpurun_out/bench_${N}gpu.log; tail -5 gpurun_out/bench_${N}gpu.err
