#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (shares, not absolutes)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    t, n = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        t[r[k]] += float(r[v].replace(",", "")) / 1e3
        n[r[k]] += 1
    tot = sum(t.values())
    if note:
        print("# " + note)
    print("# %d launches, %.1f us in total; per-launch times are cold-cache and serialised: shares, not absolutes" % (sum(n.values()), tot))
    for name in sorted(t, key=lambda x: -t[x]):
        print("%-90s n=%4d avg=%8.2f us share=%5.1f%%" % (name[:90], n[name], t[name] / n[name], 100 * t[name] / tot))


if __name__ == "__main__":
    main()
