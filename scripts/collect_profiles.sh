#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (tracked): summaries of the captures scripts/final_round.sh <tag> brought back
set -e
T=${1:-r02final}
python scripts/ncu_summary.py gpurun_out/prof_conv_$T.ncu-rep > profiles/${T}_tower_ncu_summary.txt
python scripts/ncu_summary.py gpurun_out/prof_tree_$T.ncu-rep > profiles/${T}_tree_ncu_summary.txt
python scripts/ncu_summary.py gpurun_out/prof_small_$T.ncu-rep > profiles/${T}_stem_heads_ncu_summary.txt
python scripts/launch_summary.py gpurun_out/launches_$T.csv "ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 800 --csv python bench.py --steps 1 --warmup 1 --nsims 200 --no-cpu-baseline --no-selfplay (scripts/profile_round.sh)" > profiles/${T}_launches_summary.txt
cp gpurun_out/launches_$T.csv profiles/${T}_launches.csv
cp gpurun_out/bench_$T.json profiles/${T}_bench_1gpu.json
cp gpurun_out/bench_${T}_reference.json profiles/${T}_bench_1gpu_reference_arm.json
cp gpurun_out/other_configs_$T.json profiles/${T}_other_configs.json
[ -s gpurun_out/benchmark_duels_$T.json ] && cp gpurun_out/benchmark_duels_$T.json profiles/${T}_benchmark_duels.json
cp gpurun_out/pytest_$T.log profiles/${T}_pytest_gpu.log
python - "$T" <<'P'
import json, sys
T = sys.argv[1]
rows = []
for n in (4096, 16384, 65536):
    x = json.loads(open('gpurun_out/tree_%d_%s.json' % (n, T)).read().strip().splitlines()[-1]); t = x['roofline_tree']
    rows.append(dict(trees=n, expansions_per_s=x['value'], simulations_per_s=x['simulations_per_s'], select_us_per_tick=t['select_us_per_tick'],
                     expand_backup_us_per_tick=t['expand_backup_us_per_tick'], algorithmic_GBps=t['achieved'], frac_of_measured_hbm=t['frac'], mean_depth=t['mean_depth']))
json.dump(dict(how="python bench.py --oracle-net uniform --trees N --steps 2 --warmup 1 --no-selfplay --no-cpu-baseline (tree kernels only: MCTS.RandomOracle, 600 sims per tree, Connect-Four)", rows=rows),
          open('profiles/%s_tree_pool_scaling.json' % T, 'w'), indent=1)
P
python scripts/sass_summary.py $T az_k_tower_yrow > /dev/null
mv profiles/${T}_sass_conv_yrow.txt profiles/${T}_sass_tower_yrow.txt
ls -la profiles | grep $T
