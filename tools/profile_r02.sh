#!/bin/bash
# Round-2 evidence capture (run through gpurun, ONE GPU; outputs under gpurun_out/, < 64 MiB in total).
#   tools/profile_r02.sh <tag> [launches|full3|full2|all]
# profiles/r02_* are made from these files with tools/summarise_ncu.py and tools/launch_table.py.
set -u
TAG=${1:-r02}
WHAT=${2:-all}
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-also"
K='regex:k_extend_march|k_shadow|k_shade_pre|k_shade_post|k_normals'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
if [ "$WHAT" = launches ] || [ "$WHAT" = all ]; then
  # every launch of one FULL-SIZE frame with its device time and DRAM bytes (cold-cache, serialised: compare SHARES)
  ncu --metrics $M --clock-control none -c 1100 --csv --log-file gpurun_out/${TAG}_cfg3_launches.csv \
      $B --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg3_launches.log 2>&1
  ncu --metrics $M --clock-control none -c 130 --csv --log-file gpurun_out/${TAG}_cfg2_launches.csv \
      $B --config 2 --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg2_launches.log 2>&1
fi
if [ "$WHAT" = full3 ] || [ "$WHAT" = all ]; then
  # --set full of the march / normals / shade kernels on config-3 geometry (same camera and scene, 640x360 so that ONE pass
  # holds the whole frame and every launch sees the real mix of sky / fractal / emitter lanes), depth 0
  ncu --set full --clock-control none --import-source on -k "$K" -c 5 -o gpurun_out/${TAG}_cfg3_full -f \
      $B --res 640 360 --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg3_full.log 2>&1
fi
if [ "$WHAT" = full2 ] || [ "$WHAT" = all ]; then
  ncu --set full --clock-control none --import-source on -k "$K" -c 5 -o gpurun_out/${TAG}_cfg2_full -f \
      $B --config 2 --res 512 512 --steps 1 --warmup 0 > gpurun_out/${TAG}_cfg2_full.log 2>&1
fi
ls -la gpurun_out
