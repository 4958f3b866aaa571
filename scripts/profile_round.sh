#!/bin/bash
# Round profile for profiles/: launch list of a short bench + `ncu --set full` captures of the tower conv kernel and of the
# tree kernels (select, expand+backup) at 4096 trees, late in a 600-simulation explore (deep trees).
# Usage (from the repo root, through gpurun): bash scripts/profile_round.sh <tag> [tower|tree|list|all]
set -u
tag=${1:-rXX}
what=${2:-all}
mkdir -p gpurun_out
# ncu replays kernel nodes; the device-driven WHILE graph is profiled through its host-driven twin (same kernels, same order)
export AZ_DEVICE_LOOP=0
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-selfplay"
if [ "$what" = all ] || [ "$what" = list ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 800 --csv --log-file gpurun_out/launches_${tag}.csv \
      $B --nsims 200 > gpurun_out/launches_${tag}.log 2>&1
fi
if [ "$what" = all ] || [ "$what" = tower ]; then
  # the persistent whole-tower kernel launches once per tick (AZ_TOWER=layer: az_k_conv_yrow, 14 launches per tick)
  ncu --set full --clock-control none --import-source on -k regex:az_k_tower_yrow -s 100 -c 1 -f -o gpurun_out/prof_conv_${tag} \
      $B --nsims 200 > gpurun_out/prof_conv_${tag}.log 2>&1
fi
if [ "$what" = all ] || [ "$what" = tree ]; then
  # tick ~550 of the first 600-simulation step: two launches each of select and expand+backup
  ncu --set full --clock-control none --import-source on -k "regex:az_k_select|az_k_expand_backup" -s 1100 -c 4 -f -o gpurun_out/prof_tree_${tag} \
      $B --warmup 0 --nsims 600 > gpurun_out/prof_tree_${tag}.log 2>&1
fi
ls -la gpurun_out | tail -8
