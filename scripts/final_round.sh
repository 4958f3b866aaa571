#!/bin/bash
# End-of-round evidence on one B200 (through gpurun): GPU tests, the full default bench line, the reference arm, the other
# BASELINE configs at full size, the tree kernels at growing pool sizes, and the ncu captures for profiles/.
tag=${1:-r02final}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/pytest_${tag}.log 2>&1; tail -3 gpurun_out/pytest_${tag}.log
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -2 gpurun_out/bench_${tag}.err; cut -c1-300 gpurun_out/bench_${tag}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_${tag}_reference.json
timeout 600 python scripts/run_other_configs.py > gpurun_out/other_configs_${tag}.json 2> gpurun_out/other_configs_${tag}.err; cat gpurun_out/other_configs_${tag}.json
timeout 900 python scripts/run_benchmark_duels.py > gpurun_out/benchmark_duels_${tag}.json 2> gpurun_out/benchmark_duels_${tag}.err; cat gpurun_out/benchmark_duels_${tag}.json
for T in 4096 16384 65536; do
  timeout 600 python bench.py --oracle-net uniform --trees $T --steps 2 --warmup 1 --no-selfplay --no-cpu-baseline > gpurun_out/tree_${T}_${tag}.json 2>/dev/null
done
bash scripts/profile_round.sh ${tag} all > /dev/null 2>&1
AZ_DEVICE_LOOP=0 ncu --set full --clock-control none --import-source on -k "regex:az_k_stem|az_k_heads_dense|az_k_head_conv" -s 150 -c 3 -f -o gpurun_out/prof_small_${tag} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-selfplay --nsims 100 > gpurun_out/prof_small_${tag}.log 2>&1
ls gpurun_out | tail -30
