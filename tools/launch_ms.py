#!/usr/bin/env python
"""Measurement helper (1 GPU): median per-launch ms of one bench workload's frame (event pair around every kernel, L2
flushed between frames).  python tools/launch_ms.py [workload ...]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import gravit_b200 as g  # noqa: E402
from scenes import library  # noqa: E402

g.init(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in sys.argv[1:] or ["sibenik"]:
    d = bench.WORKLOADS[name]["make"]()
    s = library.apply(d, g.GvtScene(flags=d.flags))
    e = s.engine()
    e.profile_launches(True)
    per, tot = [], []
    for i in range(13):
        flush.fill_(i)
        torch.cuda.synchronize()
        e.render_async(d.height)
        e.sync()
        if i >= 3:
            per.append([m for _, m, _ in e.launch_records()])
            tot.append(e.last_timing()[0])
    c = e.counters()
    rays = c["camera_rays"] + c["shadow_generated"] + c["secondary_generated"]
    m = statistics.median(tot)
    print(json.dumps({"workload": name, "frame_ms": round(m, 4), "Mrays_s": round(rays / m / 1e3, 1),
                      "launch_ms": [round(statistics.median(p[k] for p in per), 4) for k in range(min(len(p) for p in per))][:12]}), flush=True)
    del e, s
