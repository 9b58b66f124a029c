"""Per-kernel SASS summary of waldo_b200/libwaldo_b200.so (cuobjdump -sass): instruction count and the memory /
reduction / shuffle mnemonics that matter for the path.  usage: python profiles/sass_summary.py > profiles/r2/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "waldo_b200", "libwaldo_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "REDG", "RED.E.ADD.64", "ATOMG", "LDGSTS", "UBLKCP", "SYNCS", "HMMA", "LDS", "STS", "ATOMS", "SHFL", "REDUX", "BAR", "MUFU"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = per[cur]
        c["total"] += 1
        if op.startswith("LDG.E"):
            key = "LDG.E.128" if ".128" in op else ("LDG.E.64" if ".64" in op else "LDG.E")
            c[key] += 1
        elif op.startswith("STG.E"):
            key = "STG.E.128" if ".128" in op else ("STG.E.64" if ".64" in op else "STG.E")
            c[key] += 1
        elif op.startswith("REDG") or op.startswith("RED."):
            c["RED.E.ADD.64" if ".64" in op else "REDG"] += 1
        else:
            for k in ("ATOMG", "LDGSTS", "UBLKCP", "SYNCS", "HMMA", "LDS", "STS", "ATOMS", "SHFL", "REDUX", "BAR", "MUFU"):
                if op.startswith(k):
                    c[k] += 1
                    break


def short(name):
    try:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        d = name
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    return d


print("# SASS summary of libwaldo_b200.so (sm_100a), `cuobjdump -sass`, per kernel\n")
print("Counts are static instructions in the cubin (loads that hit the read-only path show as `LDG.E...CONSTANT`; counted with the plain ones).\n")
print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
tot = collections.Counter()
for name, c in sorted(per.items(), key=lambda kv: -kv[1]["total"]):
    print(f"| `{short(name)}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")
    tot.update(c)
print(f"| **all {len(per)} kernels** | {tot['total']} | " + " | ".join(str(tot[k]) for k in COLS) + " |")
print("\nTMA bulk copies (`UBLKCP`) with transaction barriers (`SYNCS`) and warp-level tensor-core MMAs (`HMMA`, here TF32 m16n8k8) are in the "
      "convolution kernels of f-1; tensor-map TMA (`UTMALDG` / `UTMASTG`) and tcgen05 (`UTCMMA`) occurrences: "
      + str(sum(len(re.findall(k, txt)) for k in ("UTMALDG", "UTMASTG", "UTCMMA"))) + ".")
