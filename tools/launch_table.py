"""Turns an `ncu --metrics gpu__time_duration.sum --csv` log into the per-kernel table committed under profiles/.
Usage: python tools/launch_table.py LOG.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[i_name].split("(")[0].replace("void ", "").replace("cs::", "")
    v = float(r[i_val].replace(",", ""))
    v = v / 1000.0 if r[i_unit] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.0f | %.2f | %.1f %% |" % (name, cnt, t, t / cnt, 100.0 * t / tot))
