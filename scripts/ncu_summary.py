"""Print the key metrics of each kernel in an .ncu-rep (read on the CPU box: ncu -i ... --page raw --csv)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__inst_executed.sum", "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
units = rows[1]
want = [k for k in KEYS if k in hdr]
extra = [h for h in hdr if any(s in h for s in sys.argv[2:])] if len(sys.argv) > 2 else []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:90], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in want + extra:
        print(f"   {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
