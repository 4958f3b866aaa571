"""Builds libazb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libazb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr"]
# per translation unit flags: the tree engine must not contract a*b+c (bit-exact f64 PUCT)
UNITS = {
    "az_engine.cu": ["-fmad=false"],
    "az_net.cu": [],
    "az_samples.cu": ["-fmad=false"],
    "az_comm.cu": [],
}


def _deps():
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]   # not the objects next to them
    return srcs + [os.path.join(HERE, "..", "include", "azb200.h")]


def build(force=False, verbose=False):
    objs = []
    newest_src = max(os.path.getmtime(p) for p in _deps())
    for unit, flags in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(CSRC, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest_src:
            cmd = [NVCC] + ARCH + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
