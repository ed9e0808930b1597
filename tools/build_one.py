#!/usr/bin/env python
"""Builds the BLAS of one bench mesh twice (warm-up + measured); meant to run under `ncu --metrics gpu__time_duration.sum`
for a per-kernel breakdown of the build.  python tools/build_one.py bunny|sibenik"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gravit_b200 as g  # noqa: E402
from scenes import library  # noqa: E402

g.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
m = library.bunny70k_mesh() if name == "bunny" else library.sibenik_scene(64, 36).meshes[0]
verts = np.asarray(m.verts, np.float32).reshape(-1, 3)
faces = np.asarray(m.tris, np.uint32).reshape(-1, 3) - (1 if m.one_based else 0)
for rep in range(2):
    mesh = g.Mesh(verts, faces)
    print(name, mesh.stats(), flush=True)
    mesh.close()
