#!/usr/bin/env python
"""Key counters of every kernel in an ncu report (`ncu --set full`), in the layout of the
summaries under profiles/.

  python tools/ncu_summary.py gpurun_out/r01f_c1_fused.ncu-rep
"""
import csv
import io
import subprocess
import sys

KEYS = [("duration", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs/thread", "launch__registers_per_thread"),
        ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("active threads per instruction (of 32)", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("warp instructions", "smsp__inst_executed.sum"), ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM written", "dram__bytes_write.sum"),
        ("L1 sector hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 sector hit %", "lts__t_sector_hit_rate.pct"),
        ("FMA pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("ALU pipe %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")]
STALLS = ["long_scoreboard", "wait", "branch_resolving", "short_scoreboard", "not_selected", "math_pipe_throttle",
          "no_instruction", "lg_throttle", "dispatch_stall", "mio_throttle"]


def main():
    if sys.argv[1].endswith(".csv"):     # the raw page exported on the GPU box (tools/gpu_g.sh)
        out = open(sys.argv[1]).read()
    else:
        out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        print("kernel:", d["Kernel Name"][0])
        for name, key in KEYS:
            if key in d:
                print(f"    {name:<44} {d[key][0]:>18} {d[key][1]}")
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in d:
                print(f"    stall {s + ' (warps/issue)':<38} {d[key][0]:>18}")


if __name__ == "__main__":
    main()
