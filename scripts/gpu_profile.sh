#!/bin/bash
# Run on the GPU box (under gpurun).  Produces under gpurun_out/:
#   launches_phases.csv  per-launch device times of one-launch-per-phase mode (kernel SHARES of a step)
#   launches_mega.csv    the same for the default megakernel mode
#   mega_full.ncu-rep    ncu --set full of the step megakernel (1 launch)
#   solve_full.ncu-rep   ncu --set full of the biased-solve phase kernel (3 launches)
set -u
SCENE=${1:-stack100k}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
AVN_LAUNCH_MODE=phases timeout 600 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_phases.csv \
    python bench.py --scene $SCENE --steps 1 --warmup 3 --settle 0 --no-cpu > gpurun_out/ncu_phases.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_mega.csv \
    python bench.py --scene $SCENE --steps 1 --warmup 3 --settle 0 --no-cpu > gpurun_out/ncu_mega.log 2>&1
timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k regex:step_megakernel -s 2 -c 1 -o gpurun_out/mega_full \
    python bench.py --scene $SCENE --steps 1 --warmup 3 --settle 0 --no-cpu > gpurun_out/ncu_mega_full.log 2>&1
AVN_LAUNCH_MODE=phases timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k 'regex:phase_kernel<float, \(int\)6>' -s 12 -c 3 -o gpurun_out/solve_full \
    python bench.py --scene $SCENE --steps 1 --warmup 3 --settle 0 --no-cpu > gpurun_out/ncu_solve_full.log 2>&1
ls -la gpurun_out
