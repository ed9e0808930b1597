#!/usr/bin/env python
"""One GPU: the wave schedule (k_trace / k_shade launches per bounce level) against the persistent kernel (k_persist, one
launch per frame) on the bench workloads.  python tools/persist_vs_waves.py [sibenik bunny grid gyroid[:res]] ...
Prints one JSON line per workload: ms per frame of both, rays per frame, image agreement."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import gravit_b200 as g  # noqa: E402
from gravit_b200 import schedulers  # noqa: E402
from scenes import library  # noqa: E402

names = [x for x in sys.argv[1:] if not x.startswith("--")] or ["sibenik", "bunny", "grid"]
film = next((tuple(int(v) for v in x[7:].split("x")) for x in sys.argv[1:] if x.startswith("--film=")), None)  # e.g. --film=480x270: the frame's latency chain
g.init(0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in names:
    t0 = time.perf_counter()
    if name.startswith("gyroid"):
        res = int(name.split(":")[1]) if ":" in name else library.GYROID_RES_100M
        d = library.gyroid_scene(res, 2, *(film or (3840, 2160)), brick_fn=lambda *a: library.gyroid_brick_torch(*a, device=dev))
        d.samples = next((int(x[10:]) for x in sys.argv[1:] if x.startswith("--samples=")), d.samples)  # samples^2 rays per pixel
        e, _ = schedulers.build_domain_rank(d, 0, 1)
        for m in d.meshes:
            m.verts = m.tris = None
    else:
        d = bench.WORKLOADS[name]["make"](*film) if film else bench.WORKLOADS[name]["make"]()
        s = library.apply(d, g.GvtScene(flags=d.flags))
        e = s.engine()
    build_s = time.perf_counter() - t0
    out = {"workload": name, "film": [d.width, d.height], "build_s": round(build_s, 2)}
    fbs = {}
    modes = next((x[8:].split(",") for x in sys.argv[1:] if x.startswith("--modes=")), ["waves", "persist", "persist_sp0", "persist_sp32"])
    for mode in modes:
        g.check(g.lib().gvtb200_set_tail_split(int(mode.split("_sp")[1]) if "_sp" in mode else 8))
        shade_list = next((tuple(int(v) for v in x[8:].split(",")) for x in sys.argv[1:] if x.startswith("--shade=")), (8, 14))
        for shade_at in ((8,) if mode.startswith("waves") else (shade_list if mode == "persist" else (8,))):
            if mode.startswith("persist"):
                e.domain_persist_config(0, shade_at)
            frame = (lambda: e.render_async(d.height)) if mode.startswith("waves") else (lambda: e.domain_frame_persist(0, 8))
            for _ in range(3):
                frame()
            e.sync()
            c = e.counters()
            ms = []
            for i in range(10):
                flush.fill_(i)
                torch.cuda.synchronize()
                frame()
                e.sync()
                ms.append(e.last_timing()[0])
            key = mode if mode.startswith("waves") else f"{mode}_shadesms{shade_at}"
            out[key + "_ms"] = round(statistics.median(ms), 4)
            out["rays"] = c["camera_rays"] + c["shadow_generated"] + c["secondary_generated"]
            fbs[mode.split("_")[0]] = e.fb_read(d.width, d.height)
    if "waves" in fbs and "persist" in fbs:
        out["max_abs_diff"] = float(np.abs(fbs["waves"] - fbs["persist"]).max())
    g.check(g.lib().gvtb200_set_tail_split(8))
    e.profile_counters(True)
    e.domain_frame_persist(0, 8)
    e.sync()
    out["persist_profile"] = e.domain_persist_profile()
    e.profile_counters(False)
    print(json.dumps(out), flush=True)
    del e
