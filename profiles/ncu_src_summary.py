#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output of one kernel: stall-reason totals, opcode mix (weighted by executed
instructions), and the hottest instructions / source lines.
usage: ncu -i rep --page source --csv [--print-source cuda,sass] --kernel-name regex:K | python ncu_src_summary.py [topN]"""
import csv, sys, collections, re
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line") or (len(r) > 1 and r[1] == "Source"))
print(rows[0][:2])
H = rows[hi]; idx = {h: i for i, h in enumerate(H)}
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"): break   # next launch of the same kernel
    if len(r) >= len(H) - 2: body.append(r)
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in body:
    for s in stalls: tot[s] += f(r, s)
allsamp = sum(f(r, "# Samples") for r in body)
inst = sum(f(r, "Instructions Executed") for r in body)
print(f"samples {allsamp:.0f}  warp-instructions {inst/1e6:.1f} M  static lines {len(body)}")
print("stalls: " + "  ".join(f"{k[6:]} {100*v/max(sum(tot.values()),1):.1f}%" for k, v in tot.most_common(9)))
ops = collections.Counter()
for r in body:
    src = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if m: ops[m.group(2).split(".")[0]] += f(r, "Instructions Executed")
print("opcode mix: " + "  ".join(f"{k} {100*v/max(inst,1):.1f}%" for k, v in ops.most_common(18)))
print(f"--- top {top} by samples")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
    st = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{f(r,'# Samples'):7.0f} {100*f(r,'# Samples')/max(allsamp,1):5.1f}%  exec {f(r,'Instructions Executed')/1e6:7.2f}M  {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}  | {r[idx['Source']].strip()[:110]}")
