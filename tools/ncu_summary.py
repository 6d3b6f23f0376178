#!/usr/bin/env python3
"""Key counters of one kernel from an `ncu --page raw --csv` export: python tools/ncu_summary.py file_raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
def g(key):
    for h in hdr:
        if h.endswith(key):
            return d[h]
    return None
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_issued.sum.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for k in keys:
    print("%-75s %s" % (k, g(k)))
stall = [(h, d[h]) for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [(h, d[h]) for h in hdr if "warp_issue_stalled" in h and "pct" in h]
def fl(x):
    try: return float(x)
    except Exception: return 0.0
for h, v in sorted(stall, key=lambda kv: -fl(kv[1]))[:10]:
    print("%-75s %s" % (h.split("smsp__")[-1][:75], v))
