"""Summarise `ncu -i X.ncu-rep --page raw --csv` (an `ncu --set full` capture) into a markdown table:
duration, executed warp instructions, issue-slot utilisation, FMA (integer multiply) / ALU pipe utilisation,
DRAM traffic per launch.  Usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python profiles/summarize_ncu.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %")]
cols = [(c, n) for c, n in cols if c in idx]
print("| kernel | " + " | ".join(n for _, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
    vals = []
    for c, _ in cols:
        v, u = r[idx[c]], units[idx[c]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.3g}" if f < 1e4 else f"{f:,.0f}"
        except ValueError:
            pass
        vals.append(f"{v} {u}".strip())
    print(f"| {name} | " + " | ".join(vals) + " |")
