#!/usr/bin/env python
"""Measurement helper (1 GPU): what one rank of an N-rank ImageScheduler frame costs, without the gather.

  python tools/shard_timing.py [--workload sibenik] [--worlds 1,2,4,8] [--tile-rows 8]

For every N: rank 0's row tiles of the frame (tile index 0, stride N) are rendered with an event pair around every kernel
launch; prints the frame ms, the per-launch ms and the ideal (N = 1 time / N).  Shows how much of the strong-scaling loss
is fixed per-frame cost (launches, persistent-kernel tails) rather than the exchange."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sibenik")
    ap.add_argument("--worlds", default="1,2,4,8")
    ap.add_argument("--tile-rows", type=int, default=8)
    ap.add_argument("--frames", type=int, default=10)
    a = ap.parse_args()
    import torch

    import bench
    import gravit_b200 as g
    from scenes import library

    g.init(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    d = bench.WORKLOADS[a.workload]["make"]()
    s = library.apply(d, g.GvtScene(flags=d.flags))
    e = s.engine()
    e.profile_launches(True)
    base = None
    for world in [int(x) for x in a.worlds.split(",")]:
        for _ in range(3):
            e.render_tiles(d.height, a.tile_rows, 0, world)
        e.sync()
        ms, per = [], []
        for i in range(a.frames):
            flush.fill_(i)
            torch.cuda.synchronize()
            e.render_tiles(d.height, a.tile_rows, 0, world)
            e.sync()
            ms.append(e.last_timing()[0])
            per.append([m for _, m, _ in e.launch_records()])
        m = statistics.median(ms)
        if base is None:
            base = m
        n_l = min(len(p) for p in per)
        rec = {"workload": a.workload, "world": world, "frame_ms": m, "ideal_ms": base / world,
               "launch_ms": [round(statistics.median(p[k] for p in per), 4) for k in range(n_l)],
               "sum_launch_ms": sum(statistics.median(p[k] for p in per) for k in range(n_l))}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
