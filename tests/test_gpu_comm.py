"""The two exchanges of simulate_distributed inside the C ABI (src/simulations.jl:252-290): az_samples_allgather and
az_net_broadcast over NCCL, device-resident rows.  world = 1 always runs; world = 2 needs two GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, tmp_path):
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "comm_worker.py"), str(r), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d/%d OK" % (r, world)) in out, out[-3000:]


def test_comm_single_rank(tmp_path):
    _run(1, tmp_path)


def test_comm_two_ranks(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, tmp_path)
