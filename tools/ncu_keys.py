#!/usr/bin/env python
"""The handful of ncu metrics that decide what bounds a kernel, per kernel of a report: tools/ncu_keys.py X.ncu-rep"""
import csv
import subprocess
import sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum"]
KEYS += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("-----", r[ki])
    for k in KEYS:
        if k in hdr:
            v = r[hdr.index(k)]
            try:
                if k.startswith("smsp__average") and float(v) < 0.15:
                    continue
            except ValueError:
                pass
            print("  %-88s %s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", ""), v, rows[1][hdr.index(k)]))
