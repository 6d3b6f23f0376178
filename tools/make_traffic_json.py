#!/usr/bin/env python3
"""profiles/traffic.json + a counter table from `ncu --page raw --csv` exports of the step kernel:
    python tools/make_traffic_json.py <tag> name=kernel_key ...      (reads gpurun_out/<tag>_step_<name>_raw.csv)
bench.py reads traffic.json for `roofline.traffic` (DRAM bytes per launch) and the issue-slot accounting."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "Ghz": 1e9, "Mhz": 1e6, "": 1.0, "inst": 1.0}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            out[h] = float(v) * SCALE.get(u, 1.0)
        except ValueError:
            out[h] = v
    return out


def main():
    tag = sys.argv[1]
    envs = 1 << 20
    traffic, table = {}, []
    keys = [("gpu__time_duration.sum", "duration us", 1e6), ("launch__registers_per_thread", "registers", 1), ("sm__warps_active.avg.per_cycle_active", "warps/SM", 1),
            ("smsp__inst_executed.sum", "warp-instructions M", 1e-6), ("sm__inst_issued.sum.pct_of_peak_sustained_active", "issue slots busy %", 1),
            ("sm__icc_request_hit_rate.pct", "I-cache hit %", 1), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %", 1),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %", 1), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %", 1),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %", 1), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %", 1),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1)]
    for arg in sys.argv[2:]:
        name, key = arg.split("=")
        src = os.path.join(ROOT, "gpurun_out", "%s_step_%s_raw.csv" % (tag, name))
        dst = "profiles/r2_final_step_kernel_%s_1M_steady_state_ncu_raw.csv" % name
        shutil.copy(src, os.path.join(ROOT, dst))
        d = read(src)
        dram = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        traffic[key] = {"envs": envs, "dram_bytes_per_launch": dram, "dram_read_bytes": d["dram__bytes_read.sum"], "dram_write_bytes": d["dram__bytes_write.sum"],
                        "warp_instructions_per_launch": d["smsp__inst_executed.sum"], "duration_under_ncu_us": d["gpu__time_duration.sum"] * 1e6,
                        "sm_mhz": d.get("sm__cycles_elapsed.avg.per_second", 0) / 1e6, "kernel": d["Kernel Name"], "source": dst}
        row = [name] + ["%.4g" % (d[k] * s) if isinstance(d.get(k), float) else "-" for k, _, s in keys]
        row += ["%.0f" % (dram / envs), "%.0f" % (d["smsp__inst_executed.sum"] / (envs / 32.0))]
        table.append(row)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    hdr = ["capture"] + [k[1] for k in keys] + ["DRAM B per env-step", "warp-instr per warp-step"]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for r in table:
        print("| " + " | ".join(r) + " |")


if __name__ == "__main__":
    main()
