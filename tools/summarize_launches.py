"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import collections
import sys


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    tot = sum(t for _, t in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        key = n.split("(")[0][:90]
        agg[key][0] += 1
        agg[key][1] += t
    print("launches %d, total kernel time %.2f ms" % (len(rows), tot / 1e3))
    print("%-92s %7s %10s %6s" % ("kernel", "count", "time_us", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-92s %7d %10.1f %5.1f%%" % (k, c, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
