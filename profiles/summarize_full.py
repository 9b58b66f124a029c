#!/usr/bin/env python
"""profiles/summarize_full.py <tag>: from gpurun_out/<tag>_full.ncu-rep write profiles/<tag>_ncu_full.md (one row per kernel)
and profiles/traffic.json (DRAM bytes read+written per launch, used by bench.py's `roofline.traffic`)."""
import csv, json, subprocess, sys, os
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]; idx = {h: i for i, h in enumerate(H)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__grid_size']
U = rows[1]
def gb(r, k):
    v = float(r[idx[k]]); u = U[idx[k]].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1e9)
out, lines = {}, []
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('wb_plain::', '').split('<')[0]
    if name in out: continue
    rd, wr = gb(r, 'dram__bytes_read.sum'), gb(r, 'dram__bytes_write.sum')
    out[name.replace('_async', '')] = int(rd + wr)   # bench.py's kernel names
    lines.append(f"| `{name}` | {float(r[idx[keys[0]]]):.3f} | {rd/1e9:.2f} | {wr/1e9:.2f} | {float(r[idx[keys[3]]]):.1f} | {r[idx[keys[4]]]} | "
                 f"{float(r[idx[keys[5]]]):.0f} | {float(r[idx[keys[6]]]):.0f} | {float(r[idx[keys[7]]])/1e6:.0f} | {r[idx[keys[8]]]} |")
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
md = ("| kernel | ms | DRAM read GB | DRAM write GB | DRAM % of peak | regs | warps active % | issue active % | warp-insts M | grid |\n"
      "|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(lines) + "\n")
open(os.path.join(ROOT, "profiles", tag + "_ncu_full.md"), "w").write(md)
print(md); print(out)
