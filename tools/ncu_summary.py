#!/usr/bin/env python
"""Short text summary of an `ncu --page raw --csv` export: one block per launch with the metrics the
round notes quote (duration, DRAM bytes and throughput, issue rate, occupancy, LSU wavefronts, stalls).
    python tools/ncu_summary.py raw.csv [name-filter]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
hdr, units, data = rows[0], rows[1], rows[2:]
M = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"),
     ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
     ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
     ("smsp__inst_executed.sum", "warp instructions"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("launch__registers_per_thread", "registers"), ("launch__occupancy_limit_shared_mem", "CTAs/SM by smem"),
     ("launch__occupancy_limit_registers", "CTAs/SM by regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts %"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
     ("lts__t_sector_hit_rate.pct", "L2 hit %")]
for s in ("long_scoreboard", "short_scoreboard", "barrier", "wait", "mio_throttle", "lg_throttle", "math_pipe_throttle",
          "not_selected", "branch_resolving", "no_instruction", "membar", "sleeping", "dispatch_stall", "drain", "imc_miss", "tex_throttle"):
    M.append((f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio", f"stall {s}"))
for r in data:
    name = r[hdr.index("Kernel Name")]
    if flt and flt not in name:
        continue
    print("-" * 100)
    for key, label in M:
        if key in hdr:
            i = hdr.index(key)
            print(f"  {label:24s} {r[i][:110]} {units[i]}")
