"""Key utilisation numbers + top stall reasons per launch of an .ncu-rep (reads `ncu -i ... --page raw/source --csv`)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct"]
stall = [(x, i) for i, x in enumerate(h) if x.startswith("smsp__average_warp") and "issue_stalled" in x and x.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [(x, i) for i, x in enumerate(h) if "warp_issue_stalled" in x and x.endswith(".pct")]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:70])
    for k in want:
        if k in h:
            print(f"   {k:72s} {r[h.index(k)]}")
    vals = []
    for x, i in stall:
        try:
            vals.append((float(r[i].replace(",", "")), x))
        except ValueError:
            pass
    for v, x in sorted(vals, reverse=True)[:6]:
        print(f"   stall {v:8.2f}  {x}")
