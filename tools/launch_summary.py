"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    val = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    val = val / 1e3 if unit == "ns" else val * 1e3 if unit == "ms" else val
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1
    agg[name][1] += val
    tot += val
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%9.1f us %5.1f%% n=%4d avg=%8.1f  %s" % (v[1], 100 * v[1] / tot, v[0], v[1] / v[0], k[:100]))
