"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name,
share of the total, launch count.  Usage: python tools/ncu_summary.py launches.csv [--top N]"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(dict(zip(header, r)))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
        total += v
    if "--count" in sys.argv:
        print(sum(a[1] for a in agg.values()))
        return
    print("total %.3f ms over %d launches" % (total / 1e6, sum(a[1] for a in agg.values())))
    print("%-60s %10s %7s %8s %10s" % ("kernel", "ms", "share", "launches", "avg us"))
    for name, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-60s %10.3f %6.1f%% %8d %10.1f" % (name[:60], ns / 1e6, 100 * ns / total, n, ns / n / 1e3))


if __name__ == "__main__":
    main()
