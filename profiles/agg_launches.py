"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/agg_launches.py <launches.csv> [skip_first_n_launches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
n = 0
for r in rows[start:]:
    if len(r) <= vi:
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ki]))
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{name.strip()}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f} % |")
print(f"\ntotal {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
