"""Top stall-sample source lines / opcodes of one kernel launch in an .ncu-rep."""
import csv, io, subprocess, sys
from collections import Counter
rep, skip = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "0"
def page(kind):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", kind, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
def I(x):
    try: return int(float(x))
    except Exception: return 0
rows = page("sass")
print(rows[0][:2])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
rep = next((i for i in range(1, len(data)) if data[i][0] == data[0][0]), None)  # this ncu build prints the listing twice
if rep: data = data[:rep]
tot = sum(I(r[ix["# Samples"]]) for r in data)
c = Counter(); ex = Counter()
for r in data:
    src = r[ix["Source"]].split(); op = src[0] if src else "?"
    if op.startswith("@") and len(src) > 1: op = src[1]
    c[op] += I(r[ix["# Samples"]]); ex[op] += I(r[ix["Instructions Executed"]])
print("samples", tot, "warp-instr executed", sum(ex.values()))
for op, n in c.most_common(14): print(f"  {op:32s} {100*n/max(tot,1):5.1f}%  executed {ex[op]}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
d = {s: sum(I(r[ix[s]]) for r in data) for s in stalls}
print("  stalls:", [(k, v) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:7]])
rows = page("cuda")
# find header row with "Source"
for i, r in enumerate(rows):
    if "# Samples" in r:
        hdr = r; ix = {h: i2 for i2, h in enumerate(hdr)}; data = [x for x in rows[i+1:] if len(x) == len(hdr)]; break
top = sorted(data, key=lambda r: -I(r[ix["# Samples"]]))[:14]
key = "Source" if "Source" in ix else hdr[1]
for r in top: print(f"  {I(r[ix['# Samples']]):6d}  L{r[0]:>4s} {r[ix[key]][:110]}")
