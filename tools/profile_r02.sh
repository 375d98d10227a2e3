#!/bin/bash
# Round-2 evidence capture (run through gpurun, ONE GPU).  Outputs under gpurun_out/; the summaries under profiles/ are made
# from them with tools/summarise_ncu.py.
#   $1 = tag, $2 = what: "launches" | "full3" | "full2" | "all"
set -u
TAG=${1:-r02}
WHAT=${2:-all}
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-also"
K='regex:k_extend_march|k_shadow|k_shade_pre|k_shade_post|k_normals'
if [ "$WHAT" = launches ] || [ "$WHAT" = all ]; then
  # every launch of one full-size config-3 frame with its device time (cold-cache, serialised: compare SHARES)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${TAG}_cfg3_launches.csv \
      $B --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg3_launches.log 2>&1
fi
if [ "$WHAT" = full3 ] || [ "$WHAT" = all ]; then
  # --set full of the march / normals / shade kernels on config-3 geometry (same camera and scene, 640x360 so that ONE pass
  # holds the whole frame and every launch sees the real mix of sky / fractal / emitter lanes), depths 0 and 1
  ncu --set full --clock-control none --import-source on -k "$K" -c 10 -o gpurun_out/${TAG}_cfg3_full -f \
      $B --res 640 360 --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg3_full.log 2>&1
fi
if [ "$WHAT" = full2 ] || [ "$WHAT" = all ]; then
  ncu --set full --clock-control none --import-source on -k "$K" -c 10 -o gpurun_out/${TAG}_cfg2_full -f \
      $B --config 2 --res 512 512 --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg2_full.log 2>&1
fi
ls -la gpurun_out
