mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_d.log 2>&1; tail -5 gpurun_out/pytest_d.log
timeout 600 python bench.py --no-selfplay --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -3 gpurun_out/bench_d.err; cat gpurun_out/bench_d.json
AZ_FUSED=0 timeout 600 python bench.py --no-selfplay --no-cpu-baseline > gpurun_out/bench_d_unfused.json 2>&1
NCCL_DEBUG=WARN timeout 900 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d2.json 2> gpurun_out/bench_d2.err; echo "rc=$?" >> gpurun_out/bench_d2.err; tail -30 gpurun_out/bench_d2.err; cat gpurun_out/bench_d2.json | cut -c1-600
