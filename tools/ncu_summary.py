"""Summarise an .ncu-rep (ncu --set full) into the text table kept under profiles/: one column per captured launch.

    python tools/ncu_summary.py gpurun_out/decode_kernels.ncu-rep "header line" ... > profiles/rNN_ncu_....txt
"""
import csv
import subprocess
import sys

METRICS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]

rep = sys.argv[1]
for line in sys.argv[2:]:
    print(line)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units, data = rows[0], rows[1], rows[2:]
print()
for m in METRICS:
    if m not in head:
        continue
    i = head.index(m)
    vals = [r[i] if m in ("Kernel Name", "Grid Size", "Block Size") else r[i] for r in data]
    if m == "Kernel Name":
        vals = [v.split("(")[0].replace("void ", "") for v in vals]
    print(f"{m} [{units[i]}]: " + " | ".join(vals))
