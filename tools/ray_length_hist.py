#!/usr/bin/env python
"""CPU only: how unequal are the rays of the gyroid?  Builds one gyroid mesh with the product's build bodies (tests/emul) and
counts wide-node visits + triangle tests PER RAY for the camera rays and for the shadow rays of their hits (the light sits
at the eye, so no shadow ray is occluded: an any-hit ray that never finds its hit walks everything along its path).
The longest ray of a frame is a chain of dependent steps nobody else can shorten; its length against the mean is what the
tail of a multi-GPU frame is made of.   python tools/ray_length_hist.py [res] [film_w] [film_h]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from scenes import library  # noqa: E402
from tests.scene_util import oracle_scene  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 96
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (240, 135)
d = library.gyroid_scene(res, 1, W, H)
o = oracle_scene(d)
E = C.CDLL(os.path.join(ROOT, "tests", "emul", "libgvt_emul.so"))
E.emul_build.restype = C.c_void_p
E.emul_build.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
E.emul_visit_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
nv = len(np.asarray(d.meshes[0].verts).reshape(-1, 3))
faces, _, _ = o.mesh_get(0, nv)
v = po.f32(d.meshes[0].verts).reshape(-1, 3)
f = np.ascontiguousarray(faces, np.uint32)
cam = o.camera_rays()
hits = o.intersect(0, cam, threads=16)
ok = hits["prim"] >= 0
P = cam["origin"][ok] + cam["direction"][ok] * (hits["t"][ok] * np.float32(1 - 16e-6))[:, None]
light = np.asarray(d.lights[0][0], np.float32)
shadow = np.zeros(len(P), po.RAY_DTYPE)
shadow["origin"] = P.astype(np.float32)
ld = light[None, :] - P
shadow["direction"] = (ld / np.linalg.norm(ld, axis=1)[:, None]).astype(np.float32)
shadow["t_max"] = np.float32(3.4e38)
h = C.c_void_p(E.emul_build(po._fp(v), len(v), po._fp(f), len(f)))
print(f"gyroid res {res}: {len(f)} triangles, film {W}x{H}: {len(cam)} camera rays, {int(ok.sum())} hits")
for name, rays, any_ in (("camera (closest hit)", cam, 0), ("shadow (any hit, none occluded)", shadow, 1)):
    steps = np.zeros(len(rays), np.int64)
    for i in range(len(rays)):
        out = np.zeros(8, np.uint64)
        E.emul_visit_stats(h, po._fp(rays[i:i + 1]), 1, any_, po._fp(out))
        steps[i] = int(out[0] + out[1] + out[2])  # node visits + triangle tests: the dependent steps of the ray
    s = np.sort(steps)
    q = lambda p: int(s[min(len(s) - 1, int(p * len(s)))])  # noqa: E731
    print(f"{name}: steps per ray mean {steps.mean():.1f}, median {q(0.5)}, p90 {q(0.9)}, p99 {q(0.99)}, p99.9 {q(0.999)}, max {int(s[-1])}"
          f"  -> longest / mean = {s[-1] / max(1e-9, steps.mean()):.1f}; rays above 4 x mean: {100.0 * (steps > 4 * steps.mean()).mean():.2f} %")
