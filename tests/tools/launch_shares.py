#!/usr/bin/env python
"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tests/tools/launch_shares.py gpurun_out/rNN/launches.csv "header line" > profiles/rNN_launch_shares_....txt

The list is serialised and cold-cache (ncu replays every launch on its own): compare SHARES with
bench.py's `kernel_ms`, not absolute times.
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    for line in sys.argv[2:]:
        print(line)
    print()
    rows = [r for r in csv.reader(open(path, newline="")) if r]
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = rows[head]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[head + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
        n, t = tot.get(r[ki], (0, 0.0))
        tot[r[ki]] = (n + 1, t + us)
    total = sum(t for _, t in tot.values())
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-78s n=%4d total=%10.1f us  per launch=%8.1f us  share=%5.1f%%" % (name[:78], n, t, t / n, 100 * t / total))


if __name__ == "__main__":
    main()
