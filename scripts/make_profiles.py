"""Turn the .ncu-rep captures in gpurun_out/ into the tracked summaries under profiles/ (text per capture +
profiles/ncu_traffic.json with the DRAM bytes per launch bench.py reports as `roofline.traffic`).
usage: python scripts/make_profiles.py <tag>          e.g. r1_v4"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, [dict(zip(hdr, r)) for r in rows[2:]]


traffic = {}
for name in ("ncu_gemm2", "ncu_wgrad2", "ncu_at5", "ncu_rowops"):
    rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, recs = rows_of(rep)
    U = dict(zip(hdr, units))
    lines = [f"# {name}: ncu --set full --clock-control none --import-source on (one C2 training step, B=256); per launch"]
    for d in recs:
        kn = d.get("Kernel Name", "")
        lines.append(f"== {kn[:110]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                lines.append(f"   {k:72s} {d[k]:>16s} {U.get(k, '')}")
        rd = to_bytes(d["dram__bytes_read.sum"], U["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], U["dram__bytes_write.sum"])
        dur = float(d["gpu__time_duration.sum"].replace(",", ""))
        traffic.setdefault(name, []).append({"kernel": kn[:80], "us": dur if U["gpu__time_duration.sum"] == "us" else dur,
                                             "dram_bytes_per_launch": rd + wr,
                                             "tensor_pct": float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "0") or 0)})
    open(os.path.join(ROOT, "profiles", f"{tag}_{name}.txt"), "w").write("\n".join(lines) + "\n")
    print("wrote", f"profiles/{tag}_{name}.txt", len(recs), "launches")

out = {}
# the roofline kernel of bench.py: the FFN-1 product (GELU epilogue, two bf16 outputs) at M = 16384 - the longest GELU launch
g = [t for t in traffic.get("ncu_gemm2", []) if "gemm2_kernel<0, 1, 0>" in t["kernel"] or "(bool)0, (int)1, (bool)0" in t["kernel"]]
if g:
    best = max(g, key=lambda t: t["us"])
    out["gemm2_ffn1_gelu"] = dict(best, shape_MNK=[16384, 2048, 512], algorithmic_bytes=16384 * 512 * 2 + 2048 * 512 * 2 + 2 * 16384 * 2048 * 2)
w = traffic.get("ncu_wgrad2", [])
if w:
    out["wgrad2_group_layer"] = max(w, key=lambda t: t["us"])
for k in ("ncu_at5", "ncu_rowops"):
    for t in traffic.get(k, []):
        out.setdefault(f"{k}:{t['kernel'][:40]}", t)
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
