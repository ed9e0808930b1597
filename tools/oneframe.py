#!/usr/bin/env python
"""Renders a few frames of one bench workload (target for ncu captures).  python tools/oneframe.py sibenik [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gravit_b200 as g  # noqa: E402
from scenes import library  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sibenik"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g.init(0)
d = library.sibenik_scene(3840, 2160, depth=5, samples=3) if name == "incoherent" else bench.WORKLOADS[name]["make"]()
s = library.apply(d, g.GvtScene(flags=d.flags))
e = s.engine()
for _ in range(frames):
    e.render_async(d.height)
e.sync()
print(e.counters())
