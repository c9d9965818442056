"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg.setdefault(row["Kernel Name"][:90], []).append(v)
tot = sum(sum(v) for v in agg.values())
print("%-92s %6s %12s %10s %6s" % ("kernel", "n", "total_us", "avg_us", "%"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
    print("%-92s %6d %12.1f %10.1f %6.1f" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
print("TOTAL us %.1f over %d launches" % (tot, sum(len(v) for v in agg.values())))
