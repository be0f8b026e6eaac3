"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for x in csv.DictReader(lines):
        name = x["Kernel Name"]
        name = name.replace("void ", "").replace("mmer::", "")
        cut = name.find("(")
        if cut > 0:
            name = name[:cut]
        v = float(x["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(x["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{'total us':>10s} {'n':>5s} {'mean us':>9s} {'share':>6s}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} {c:5d} {t / c:9.1f} {100 * t / tot:5.1f}%  {n[:90]}")
    print(f"{tot:10.1f} us in {sum(c for c, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
