#!/bin/bash
# Run on the GPU box (under gpurun): ncu evidence for the current tree.  Produces under gpurun_out/ (TAG = first argument, default r02):
#   TAG_launches_<scene>.csv      per-launch device times of the default mode (kernel SHARES of a step)
#   TAG_mega_<scene>.ncu-rep      ncu --set full of the step megakernel (1 launch)  + TAG_mega_<scene>_raw.csv (--page raw)
# Scenes: stack100k (wavefront records, f32) and spheres1m (barrier schedule, f64, MAXP = 1).
set -u
TAG=${1:-r02}
shift || true
SCENES=${@:-stack100k spheres1m}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for SCENE in $SCENES; do
  ARGS="bench.py --scene $SCENE --steps 2 --warmup 3 --no-cpu --no-partition --no-pass"
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_launches_${SCENE}.csv python $ARGS > gpurun_out/${TAG}_ncu_launches_${SCENE}.log 2>&1
  timeout 1200 $NCU --set full --import-source on --kernel-name-base demangled -k regex:step_megakernel -s 3 -c 1 -f -o gpurun_out/${TAG}_mega_${SCENE} \
      python $ARGS > gpurun_out/${TAG}_ncu_full_${SCENE}.log 2>&1
  ncu -i gpurun_out/${TAG}_mega_${SCENE}.ncu-rep --page raw --csv > gpurun_out/${TAG}_mega_${SCENE}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_mega_${SCENE}.ncu-rep --page details > gpurun_out/${TAG}_mega_${SCENE}_details.txt 2>/dev/null
done
ls -la gpurun_out | tail -20
