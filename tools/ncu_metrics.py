#!/usr/bin/env python
"""Per-launch ncu counters of one frame of a bench workload, written as the JSON bench.py reads for its roofline
(issue term from smsp__inst_executed, HBM term from dram__bytes, SIMT efficiency and L1 wavefronts beside them).

  on the GPU box:   python tools/ncu_metrics.py run sibenik [waves|persist]     -> gpurun_out/ncu_metrics_<workload>_<mode>.json
  here:             cp gpurun_out/ncu_metrics_*.json profiles/r2_ncu_metrics_*.json

The JSON carries the hash of gravit_b200/csrc at capture time; bench.py refuses a capture whose hash is not the hash of the
sources it runs (a stale capture says nothing about the current kernels)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
UNIT = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "hz": 1.0, "khz": 1e3, "mhz": 1e6, "ghz": 1e9, "cycle/second": 1.0,
        "cycle/nsecond": 1e9, "cycle/usecond": 1e6}


def parse(csv_path):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    hdr = rows[0]
    ci = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = {}
    for r in rows[1:]:
        if len(r) <= ci["Metric Value"]:
            continue
        lid = int(r[ci["ID"]])
        d = launches.setdefault(lid, {"kernel": r[ci["Kernel Name"]].split("(")[0]})
        try:
            v = float(r[ci["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        d[r[ci["Metric Name"]]] = v * UNIT.get(r[ci["Metric Unit"]].lower(), 1.0)
    return [launches[k] for k in sorted(launches)]


def main():
    if sys.argv[1] == "run":
        import gravit_b200 as g

        workload = sys.argv[2]
        mode = sys.argv[3] if len(sys.argv) > 3 else "waves"
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        raw = os.path.join(ROOT, "gpurun_out", f"ncu_metrics_{workload}_{mode}.csv")
        # one warm-up frame (skipped: -s is found by counting the frame's launches in a first, unprofiled pass), one profiled
        cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", "regex:k_trace|k_shade|k_persist|k_camera_route", "--csv", "--log-file", raw,
               sys.executable, os.path.join(ROOT, "tools", "oneframe.py"), workload, "2", mode]
        subprocess.check_call(cmd, cwd=ROOT)
        launches = parse(raw)
        half = len(launches) // 2  # two identical frames: keep the second (warm)
        out = {"workload": workload, "mode": mode, "source_hash": g.source_hash(), "launches": launches[half:],
               "how": "ncu --metrics ... --clock-control none, second of two frames of tools/oneframe.py (cold-ish caches, serialised launches)"}
        dst = os.path.join(ROOT, "gpurun_out", f"ncu_metrics_{workload}_{mode}.json")
        json.dump(out, open(dst, "w"), indent=1)
        for i, l in enumerate(out["launches"]):
            print(i, l["kernel"], {k: l.get(k) for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum")})
    else:
        print(json.dumps(parse(sys.argv[1]), indent=1))


if __name__ == "__main__":
    main()
