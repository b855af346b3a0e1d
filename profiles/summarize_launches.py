"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"][:64]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
        n += 1
    tot = sum(d[1] for d in agg.values())
    print(f"launches {n}  total {tot:.0f} us")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:66s} n={c:4d} total={t:9.1f}us avg={t / c:8.1f}us share={t / tot * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
