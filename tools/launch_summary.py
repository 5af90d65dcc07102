"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        t = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        t *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = row["Kernel Name"]
        short = re.sub(r"\(.*", "", name).replace("void cvxb::<unnamed>::", "").replace("cvxb::<unnamed>::", "")
        rows.append((short, t, row.get("Grid Size", "")))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, t, g in rows:
        agg[s][0] += 1
        agg[s][1] += t
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f%% | %.1f |" % (k, v[0], v[1] / 1e6, v[1] / tot * 100, v[1] / v[0] / 1e3))
    print("\ntotal %.3f ms over %d launches" % (tot / 1e6, len(rows)))


if __name__ == "__main__":
    main()
