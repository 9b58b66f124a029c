#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. usage: launch_summary.py file.csv [steps]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    d = dict(zip(H, r))
    if 'gpu__time_duration.sum' not in d.get('Metric Name', ''): continue
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    ms = v / 1e6 if u.startswith('n') else v / 1e3 if u.startswith('u') else v
    k = d['Kernel Name'][:64]
    agg[k][0] += 1; agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches/step | avg ms | ms/step | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]/steps:g} | {v[1]/v[0]:.3f} | {v[1]/steps:.3f} | {100*v[1]/tot:.1f}% |")
print(f"| total | | | {tot/steps:.3f} | |")
