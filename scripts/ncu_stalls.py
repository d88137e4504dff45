"""Top stall sites of one kernel in an .ncu-rep (source page, SASS view).
usage: python scripts/ncu_stalls.py rep.ncu-rep <kernel regex> [launch-skip] [top-n]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:150])
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[1] == "Source":
        break
    data.append(r)
iS, isrc = hdr.index("# Samples"), hdr.index("Source")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print("instructions", len(data), "total samples", tot)
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print("overall:", [(k, round(100 * v / max(tot, 1), 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:topn]
for i in sorted(top):
    r = data[i]
    st = {hdr[j][6:]: int(r[j]) for j in stalls if int(r[j]) > 0}
    print(f"{i:5d} {100 * int(r[iS]) / max(tot, 1):5.1f}%  {r[isrc].strip()[:80]:80s}", sorted(st.items(), key=lambda kv: -kv[1])[:3])
