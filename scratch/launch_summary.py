#!/usr/bin/env python
"""Per-kernel totals of ncu launch lists (--metrics gpu__time_duration.sum --csv): python scratch/launch_summary.py a.csv [b.csv ...]"""
import collections, csv, re, sys

for path in sys.argv[1:]:
    try:
        lines = [l for l in open(path) if not l.startswith("==")]
    except OSError as ex:
        print(f"## {path}: {ex}"); continue
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values()) or 1.0
    print(f"## {path}: {sum(cnt.values())} launches, {T / 1e6:.2f} ms summed device time (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':90s} {'launches':>9s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
        print(f"{k[:90]:90s} {cnt[k]:9d} {v / 1e6:10.3f} {v / T:7.1%} {v / cnt[k] / 1e3:10.1f}")
