"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel name the
number of launches and the total / last duration in microseconds.  usage: launch_summary.py
file.csv [last_n]  (last_n: only the last n launches, e.g. those of the final repetition)"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ix = {k: i for i, k in enumerate(hdr)}
for row in r:
    if len(row) < len(hdr) or row[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(row[ix["Metric Value"]].replace(",", ""))
    unit = row[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
    rows.append((row[ix["Kernel Name"]], v))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for name, v in rows:
    name = name.split("(")[0][:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-90s %6s %12s" % ("kernel", "n", "us"))
for name, (n, v) in agg.items():
    print("%-90s %6d %12.1f" % (name, n, v))
print("%-90s %6d %12.1f" % ("total", len(rows), tot))
