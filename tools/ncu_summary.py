#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the text table committed under profiles/ and the JSON bench.py reads for
roofline.traffic.  python tools/ncu_summary.py raw.csv out.txt [out.json workload launch_index]"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
name_i = hdr.index("Kernel Name")
out = open(sys.argv[2], "w")
out.write("# ncu --set full --clock-control none; one column per captured launch, in launch order\n")
out.write("# kernels: " + " | ".join(r[name_i][:40] for r in data) + "\n")
for k in KEEP:
    if k in hdr:
        i = hdr.index(k)
        out.write(f"{k:90s} {units[i]:16s} " + " | ".join(r[i] for r in data) + "\n")
out.close()
if len(sys.argv) > 5:
    li = int(sys.argv[5])
    def val(k):
        i = hdr.index(k)
        v = float(data[li][i].replace(",", ""))
        u = units[i].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return v * mult
    j = {"workload": sys.argv[4], "launch_index_in_capture": li, "kernel": data[li][name_i],
         "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
         "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
         "gpu_time_us": float(data[li][hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(units[hdr.index("gpu__time_duration.sum")], 1),
         "source": "ncu --set full --clock-control none, tools/oneframe.py sibenik"}
    json.dump(j, open(sys.argv[3], "w"), indent=1)
