#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel histogram of the Blackwell-specific mnemonics (tcgen05 -> UTC*MMA, tcgen05.ld ->
LDTM, TMA -> UTMALDG/UTMASTG, legacy HMMA must be absent) + the full listing of the dominant kernel.
Usage: python scripts/sass_summary.py <tag>   (reads alphazero.jl_b200/csrc/*.o, writes profiles/<tag>_sass_*.txt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOM", "SYNCS", "ELECT", "HMMA", "DMMA", "DFMA", "DADD", "DMUL",
        "LDG", "STG", "ATOMG", "RED", "SHFL", "LDS", "STS"]


def main():
    tag = sys.argv[1]
    full = sys.argv[2] if len(sys.argv) > 2 else "az_k_conv_yrowILi1E"
    out = []
    listing = None
    for obj in ("az_net.o", "az_engine.o", "az_samples.o", "az_comm.o"):
        p = os.path.join(ROOT, "alphazero.jl_b200", "csrc", obj)
        if not os.path.exists(p):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True).stdout
        demangle = {}
        cur, body = None, collections.defaultdict(list)
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                continue
            if cur:
                body[cur].append(line)
        for fn, lines in body.items():
            cnt = collections.Counter()
            n = 0
            for l in lines:
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
                if not m:
                    continue
                n += 1
                op = m.group(1)
                for k in KEYS:
                    if op.startswith(k):
                        cnt[k] += 1
                        break
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:100]
            out.append("%-8s %-100s insts %6d  %s" % (obj, name, n, " ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k])))
            if full in fn:
                listing = (name, lines)
    with open(os.path.join(ROOT, "profiles", tag + "_sass_mnemonics.txt"), "w") as f:
        f.write("# cuobjdump -sass of the objects linked into libazb200.so (sm_100a): instructions per kernel and counts of the\n"
                "# mnemonics that matter (UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store,\n"
                "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA (legacy mma.sync) must not appear)\n")
        f.write("\n".join(out) + "\n")
    if listing:
        with open(os.path.join(ROOT, "profiles", tag + "_sass_conv_yrow.txt"), "w") as f:
            f.write("# cuobjdump -sass: %s\n" % listing[0])
            for l in listing[1]:
                m = re.search(r"(/\*[0-9a-f]{4}\*/\s+.*?;)", l)
                if m:
                    f.write(m.group(1) + "\n")
    print("\n".join(o for o in out if "conv_yrow" in o or "select" in o or "gemm_tc" in o))


if __name__ == "__main__":
    main()
