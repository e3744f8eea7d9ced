"""Summarise a .ncu-rep (read here, on the CPU box) into the few lines the roofline discussion needs."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_red.sum"]
for v in rows[2:]:
    print("-" * 100)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print("%-70s %-12s %s" % (k, units[i], v[i]))
    st = [(float(v[i].replace(",", "")), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    tot = sum(s for s, _ in st)
    print("warp stall reasons (share of stalled warp-cycles per issue):")
    for s, h in sorted(st, reverse=True)[:6]:
        print("   %-30s %6.1f%%" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), 100 * s / tot))
