#!/usr/bin/env python
"""Renders a few frames of one bench workload (target for ncu captures).  python tools/oneframe.py sibenik [frames] [waves|persist]
gyroid[:res] builds the configs[3] isosurface (all bricks on this GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gravit_b200 as g  # noqa: E402
from scenes import library  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sibenik"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "waves"
g.init(0)
if name.startswith("gyroid"):
    import torch

    from gravit_b200 import schedulers

    res = int(name.split(":")[1]) if ":" in name else library.GYROID_RES_100M
    d = library.gyroid_scene(res, 2, 3840, 2160, brick_fn=lambda *a: library.gyroid_brick_torch(*a, device=torch.device("cuda", 0)))
    e, _ = schedulers.build_domain_rank(d, 0, 1)
else:
    d = library.sibenik_scene(3840, 2160, depth=5, samples=3) if name == "incoherent" else bench.WORKLOADS[name]["make"]()
    s = library.apply(d, g.GvtScene(flags=d.flags))
    e = s.engine()
for _ in range(frames):
    if mode == "persist":
        e.domain_frame_persist(0, 8)
    else:
        e.render_async(d.height)
e.sync()
print(e.counters())
