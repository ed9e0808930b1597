#!/usr/bin/env python
"""Measurement helper (run on the GPU box): frame time of the bench workloads for a list of trace-kernel tunings.

  python tools/sweep.py [--tunes "8,12,2;32,32,0"] [--workloads sibenik,bunny,grid] [--out gpurun_out/sweep.jsonl]

Every line: workload, tuning, median frame ms (CUDA events, L2 flushed between frames), Mrays/s, per-wave ms, counters.
The framebuffer checksum must not depend on the tuning (scheduling knobs only)."""
import argparse
import json
import os
import statistics
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tunes", default="8,12,2")
    ap.add_argument("--workloads", default="sibenik,bunny,grid")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch

    import bench
    import gravit_b200 as g
    from scenes import library

    g.init(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    out = open(a.out, "a")
    tunes = [tuple(int(x) for x in t.split(",")) for t in a.tunes.split(";") if t]
    for name in a.workloads.split(","):
        if name == "incoherent":
            d = library.sibenik_scene(3840, 2160, depth=5, samples=3)
        else:
            d = bench.WORKLOADS[name]["make"]()
        s = library.apply(d, g.GvtScene(flags=d.flags))
        e = s.engine()
        ref_crc = None
        for tune in tunes:
            g.check(g.lib().gvtb200_set_tuning(*(tune + (1,) * (4 - len(tune)))))
            for _ in range(3):
                e.render_async(d.height)
            e.sync()
            c = e.counters()
            rays = c["camera_rays"] + c["shadow_generated"] + c["secondary_generated"]
            ms, waves = [], []
            for i in range(a.frames):
                flush.fill_(i)
                torch.cuda.synchronize()
                e.render_async(d.height)
                e.sync()
                t = e.last_timing()
                ms.append(t[0])
                waves.append(t[1])
            fb = e.fb_read(d.width, d.height)
            crc = zlib.crc32(np.ascontiguousarray(fb).tobytes()) if d.depth == 1 and len(d.lights) == 1 and d.samples == 1 else None
            if ref_crc is None:
                ref_crc = crc
            m = statistics.median(ms)
            rec = {
                "tag": a.tag, "workload": name, "tune": tune, "ms": m, "min_ms": min(ms), "Mrays_s": rays / m / 1e3, "rays": rays,
                "last_wave_ms": [statistics.median(w[i] for w in waves) for i in range(3)],
                "stack_overflows": c["stack_overflows"], "fb_crc_same": (crc == ref_crc), "hits": c["primary_hits"], "fb_adds": c["fb_adds"],
            }
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            out.flush()
        del e, s


if __name__ == "__main__":
    main()
