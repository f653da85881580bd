#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel."""
import csv, sys, collections, re
def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        agg[name][0] += 1; agg[name][1] += us; total += us; n += 1
    print(f"{n} launches, {total:.1f} us total device time")
    print("| share | launches | time us | avg us | kernel |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {100*t/total:.2f}% | {c} | {t:.1f} | {t/c:.2f} | `{k}` |")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
