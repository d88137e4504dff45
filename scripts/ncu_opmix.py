"""Instruction mix of a kernel from `ncu -i X.ncu-rep --page source --csv`: warp instructions executed per SASS opcode,
and the top stall-sample lines.  usage: python scripts/ncu_opmix.py <rep> [tasks]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
tasks = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
h = rows[1]
si, ie, ss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
c, st = Counter(), []
tot = 0
for r in rows[2:]:
    try:
        n = int(r[ie])
    except Exception:
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    c[op.split(".")[0]] += n
    tot += n
    st.append((int(r[ss] or 0), r[si].strip()[:90]))
print(f"total warp instructions {tot}  per task {tot / tasks:.0f}")
for k, v in c.most_common(18):
    print(f"  {k:10s} {v:11d} {100 * v / tot:5.1f}%")
print("top stall lines:")
tots = sum(s for s, _ in st) or 1
for s, l in sorted(st, reverse=True)[:12]:
    print(f"  {100 * s / tots:5.1f}%  {l}")
