"""The C ABI driven by a plain-C host on the GPU (examples/selfplay_c_abi.c): no Python, no torch in the process -- the
library alone creates the context, runs self-play to completion, fetches the samples and does the replay-buffer
operations.  (The CPU suite compiles the same program and checks its loud failure without a GPU.)"""
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_host_runs_selfplay_on_the_gpu(tmp_path):
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("gcc not available")
    libdir = os.path.join(ROOT, "alphazero.jl_b200")
    exe = str(tmp_path / "selfplay_c")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "selfplay_c_abi.c"), "-L" + libdir, "-lazb200", "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe, "24", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"games (\d+) samples (\d+)\s+simulations (\d+) expansions (\d+)", r.stdout)
    assert m and int(m.group(1)) == 24 and int(m.group(2)) >= 24 * 7 and int(m.group(4)) > 0, r.stdout
    m2 = re.search(r"augmented (\d+) -> distinct states (\d+)", r.stdout)
    assert m2 and int(m2.group(1)) == 2 * int(m.group(2)) and 0 < int(m2.group(2)) <= int(m2.group(1)), r.stdout
    # the process loaded libazb200.so and nothing from Python / torch
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libazb200.so" in out and "libtorch" not in out and "libpython" not in out
