#!/usr/bin/env python
"""BLAS build time / size for the bench meshes and a large gyroid brick.  python tools/buildtime.py [res] [only] [reps]
`only`: just the gyroid brick (the target of an ncu launch list of one big build)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gravit_b200 as g  # noqa: E402
from scenes import library  # noqa: E402

g.init(0)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = len(sys.argv) > 2 and sys.argv[2] == "only"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cases = [] if only else [("bunny70k", library.bunny70k_mesh()), ("sibenik", library.sibenik_scene(64, 36).meshes[0])]
t0 = time.time()
v, t = library.gyroid_brick(res, (0, 0, 0), 1)
print(f"gyroid res {res}: {len(t)} tris generated in {time.time()-t0:.1f} s", flush=True)
cases.append((f"gyroid{res}", library.MeshDesc(v, t, True, None)))
for name, m in cases:
    verts = np.asarray(m.verts, np.float32).reshape(-1, 3)
    faces = np.asarray(m.tris, np.uint32).reshape(-1, 3) - (1 if m.one_based else 0)
    us, walls = [], []
    for rep in range(reps):
        t0 = time.time()
        mesh = g.Mesh(verts, faces)
        walls.append((time.time() - t0) * 1e3)
        st = mesh.stats()
        us.append(st["build_us"])
        mesh.close()
    warm = sorted(us[1:]) or us
    print(name, "tris", len(faces), {k: st[k] for k in ("wide_nodes", "depth")}, f"build_us first {us[0]} warm min {warm[0]} median {warm[len(warm)//2]}",
          f"-> {len(faces) / warm[len(warm)//2]:.1f} Mtris/s; wall median {sorted(walls)[len(walls)//2]:.1f} ms", flush=True)
