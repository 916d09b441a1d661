"""Summarise an .ncu-rep (read on the CPU box) into a small text file for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none, summary of %s\n" % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write("\n")
        for k in KEYS:
            if k in d:
                f.write("%-90s %s %s\n" % (k, d[k], units[hdr.index(k)]))
        try:   # ncu prints one unit per COLUMN (first result's scale); values are already in that unit
            sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"]) * sc[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(d["dram__bytes_write.sum"]) * sc[units[hdr.index("dram__bytes_write.sum")]]
            f.write("%-90s %.4f Gbyte\n" % ("dram traffic (read+write)", (rd + wr) / 1e9))
        except (KeyError, ValueError):
            pass
print(open(out).read())
