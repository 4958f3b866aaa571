#!/bin/bash
# one GPU box visit: GPU test-suite, the default bench, optional extras (args: tag [extra commands...])
mkdir -p gpurun_out
tag=${1:-x}
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_${tag}.log 2>&1; tail -4 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --no-selfplay --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -3 gpurun_out/bench_${tag}.err; cut -c1-400 gpurun_out/bench_${tag}.json
