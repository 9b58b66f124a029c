#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing of one kernel with `nvdisasm -g -c` line info of the same build and
aggregate stall samples / executed instructions per source line.
usage: ncu_lines.py <ncu_src.csv> <nvdisasm.txt> <mangled-kernel-substring> [topN]"""
import csv, sys, re, collections
src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
# --- line table of the kernel
line_of = {}
cur = None; inside = False
for ln in open(dis, errors="ignore"):
    if ln.startswith("//---------------------"):
        inside = (".text." in ln and kern in ln)
        cur = None
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if len(r) > 1 and r[1] == "Source")
H = rows[hi]; idx = {h: i for i, h in enumerate(H)}
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"): break   # next launch of the same kernel
    body.append(r)
base = int(body[0][0], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
seen = set()
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
for r in body:
    if len(r) < len(H) - 2: continue
    off = int(r[0], 16) - base
    if off in seen: continue
    seen.add(off)
    key = line_of.get(off)
    a = agg[key]
    a[0] += float(r[idx["# Samples"]] or 0); a[1] += float(r[idx["Instructions Executed"]] or 0)
    for s in stalls:
        a[2][s[6:]] += float(r[idx[s]] or 0)
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f"{kern}: samples {ts:.0f}, warp-instructions {ti/1e6:.1f} M")
srcs = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        import os
        for d in ("/root/repo/waldo_b200/csrc/", "/root/repo/include/"):
            p = d + k[0]
            if os.path.exists(p):
                srcs.setdefault(p, open(p).read().splitlines())
                text = srcs[p][k[1] - 1].strip()[:90]
    st = ", ".join(f"{n}:{100*v/max(a[0],1):.0f}%" for n, v in a[2].most_common(2))
    print(f"{100*a[0]/ts:5.1f}% samp {100*a[1]/ti:5.1f}% inst  {str(k[0])+':'+str(k[1]) if k else '?':28s} [{st}] {text}")
