#!/bin/bash
# Tower-kernel profile for profiles/: launch list of a short bench + one `ncu --set full` capture of conv1/conv2.
# Usage (from the repo root, through gpurun): bash scripts/profile_tower.sh <tag>
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 1 --nsims 100 --no-cpu-baseline --no-selfplay > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:az_k_conv_c4_2sm -s 120 -c 2 -f -o gpurun_out/prof_conv_${tag} \
    python bench.py --steps 1 --warmup 1 --nsims 100 --no-cpu-baseline --no-selfplay > gpurun_out/prof_conv_${tag}.log 2>&1
ls -la gpurun_out | tail -5
