"""Print a compact summary of an .ncu-rep (raw page) -- used to write profiles/*.txt."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active",
                        "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
                        "smsp__issue_active.avg.pct", "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled", "sm__throughput.avg.pct",
                        "gpu__dram_throughput.avg.pct", "lts__t_sector_hit_rate", "sm__inst_executed_pipe", "smsp__inst_executed.sum ",
                        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, units, rows = r[0], r[1], r[2:]
for i, name in enumerate(h):
    if any(p in name for p in pats) and "dshared" not in name:
        vals = [row[i] for row in rows]
        if all(v in ("0", "", "n/a") for v in vals):
            continue
        print(f"{name} [{units[i]}]: {', '.join(vals[:3])}")
