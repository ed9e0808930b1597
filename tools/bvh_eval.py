#!/usr/bin/env python
"""CPU evaluation of BLAS quality (no GPU): builds the BLAS with the product's build bodies (tests/emul) and counts wide nodes
visited (inner / bottom) and triangles tested per ray for camera rays, shadow rays and diffuse bounce rays of a scene, plus
the fill histogram of the wide nodes.  The counts are what the trace kernel's COUNT instantiation reports on the GPU
(sibenik: 13.1 nodes + 8.5 triangles per bounce ray on both), so build heuristics can be compared here before a GPU run:
    python tools/bvh_eval.py sibenik          # rebuild tests/emul/libgvt_emul.so with other -D flags to compare variants"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from scenes import library  # noqa: E402
from tests.scene_util import oracle_scene  # noqa: E402


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "sibenik"
    d = {"sibenik": lambda: library.sibenik_scene(320, 180), "bunny": lambda: library.bunny_scene(320, 180),
         "simple": lambda: library.simple_scene(128, 128)}[scene]()
    o = oracle_scene(d)
    E = C.CDLL(os.environ.get("GVT_EMUL_SO", os.path.join(ROOT, "tests", "emul", "libgvt_emul.so")))
    E.emul_build.restype = C.c_void_p
    E.emul_build.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    E.emul_destroy.argtypes = [C.c_void_p]
    E.emul_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    E.emul_stats.argtypes = [C.c_void_p] * 2
    k = 0
    nv = len(np.asarray(d.meshes[k].verts).reshape(-1, 3))
    faces, _, bbox = o.mesh_get(k, nv)
    v = po.f32(d.meshes[k].verts).reshape(-1, 3)
    f = np.ascontiguousarray(faces, np.uint32)
    cam = o.camera_rays()
    hits = o.intersect(k, cam, threads=8)
    ok = hits["prim"] >= 0
    rng = np.random.default_rng(7)
    P = cam["origin"][ok] + cam["direction"][ok] * (hits["t"][ok] * np.float32(1 - 16e-6))[:, None]
    tri = v[f[hits["prim"][ok]]]
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.maximum(np.linalg.norm(n, axis=1), 1e-30)[:, None]
    n[np.einsum("ij,ij->i", n, cam["direction"][ok]) > 0] *= -1
    # cosine-weighted directions about n
    u1, u2 = rng.random(len(n)), rng.random(len(n))
    r, phi = np.sqrt(u1), 2 * np.pi * u2
    a = np.where(np.abs(n[:, :1]) < 0.5, [[1.0, 0, 0]], [[0, 1.0, 0]])
    t1 = np.cross(n, a)
    t1 /= np.linalg.norm(t1, axis=1)[:, None]
    t2 = np.cross(n, t1)
    dirs = t1 * (r * np.cos(phi))[:, None] + t2 * (r * np.sin(phi))[:, None] + n * np.sqrt(1 - u1)[:, None]
    bounce = np.zeros(len(n), po.RAY_DTYPE)
    bounce["origin"], bounce["direction"] = P.astype(np.float32), dirs.astype(np.float32)
    bounce["t_max"] = np.float32(3.4e38)
    light = np.asarray(d.lights[0][0], np.float32)
    shadow = np.zeros(len(n), po.RAY_DTYPE)
    shadow["origin"] = P.astype(np.float32)
    ld = light[None, :] - P
    shadow["direction"] = (ld / np.linalg.norm(ld, axis=1)[:, None]).astype(np.float32)
    shadow["t_max"] = np.float32(3.4e38)
    sets = [("camera", cam, 0), ("shadow(any)", shadow, 1), ("bounce", bounce, 0)]
    E.emul_visit_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    E.emul_node_hist.argtypes = [C.c_void_p] * 2
    E.emul_ideal_visits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    print(f"{scene}: {len(f)} faces; rays: " + ", ".join(f"{nm} {len(r_)}" for nm, r_, _ in sets))
    h = C.c_void_p(E.emul_build(po._fp(v), len(v), po._fp(f), len(f)))
    st = np.zeros(4, np.uint64)
    E.emul_stats(h, po._fp(st))
    hist = np.zeros(14, np.uint64)
    E.emul_node_hist(h, po._fp(hist))
    print(f"wide nodes {int(st[0])} (bottom {int(hist[13])}), triangle refs {int(st[1])}, depth {int(st[2])} (binary tree {int(st[3])} levels); children per node 0..8: {hist[:9].tolist()}; "
          f"triangles per leaf slot 1..3: {hist[10:13].tolist()}")
    total = 0.0
    for nm, rays, any_ in sets:
        out = np.zeros(8, np.uint64)
        E.emul_visit_stats(h, po._fp(rays), len(rays), any_, po._fp(out))
        inner, bottom, tt = (float(x) / len(rays) for x in out[:3])
        top = [float(x) / max(1.0, float(out[0] + out[1])) for x in out[3:7]]
        cost = 8 * (inner + bottom) + 3 * tt
        total += cost
        ideal = f"; visits that hit nothing inside: {100.0 * float(out[7]) / max(1.0, float(out[0] + out[1])):.1f}%"
        if not any_:
            io = np.zeros(2, np.uint64)
            E.emul_ideal_visits(h, po._fp(rays), len(rays), po._fp(io))
            ideal += f"; nearest-first with distance culling would need {float(io[0]) / len(rays):.2f} nodes, {float(io[1]) / len(rays):.2f} triangles"
        print(f"  {nm:12s}: {inner + bottom:6.2f} nodes ({inner:5.2f} inner + {bottom:5.2f} bottom), {tt:5.2f} triangles, L1 wavefronts {cost:6.1f}; visits to the first 9/73/200/585 nodes: " + "/".join(f"{100 * x:.0f}%" for x in top) + ideal)
    print(f"  total wavefronts per ray triple: {total:6.1f}")
    E.emul_destroy(h)


if __name__ == "__main__":
    main()
