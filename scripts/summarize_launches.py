"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path: str) -> None:
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(row["Metric Unit"], v)
        name = re.sub(r"<.*", "", row["Kernel Name"]).replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {t / tot * 100:.1f}% |")
    print(f"\ntotal {tot:.1f} us over {sum(c for c, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
