#!/usr/bin/env python
"""Writes tests/golden/readers/readers_golden.json: SHA-256 digests of what the REAL third-party readers the reference
links (tinyobjloader, ply.c -- compiled from /root/reference by `make -C oracle ref`) produce for
  * the committed fixtures in tests/golden/readers/ (so the check travels to hosts without /root/reference), and
  * the reference's own scene files under /root/reference/data (checked only where that tree exists).
Run in the build container:  python tests/golden/make_reader_golden.py"""
import ctypes as C
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def ref_lib():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgvt_ref_readers.so"))
    vp = C.c_void_p
    L.gvtref_obj_load.argtypes = [C.c_char_p, vp]
    L.gvtref_obj_get.argtypes = [vp] * 5
    L.gvtref_ply_load.argtypes = [C.c_char_p, vp]
    L.gvtref_ply_get.argtypes = [vp] * 3
    return L


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_obj(L, path):
    n = np.zeros(4, np.uint32)
    assert L.gvtref_obj_load(path.encode(), p(n)) == 0, path
    v, f, fm = np.zeros((n[0], 3), np.float32), np.zeros((n[1], 3), np.uint32), np.zeros(n[1], np.int32)
    m, nr = np.zeros((n[2], 9), np.float32), np.zeros((n[3], 3), np.float32)
    L.gvtref_obj_get(p(v), p(f), p(fm), p(m), p(nr))
    return dict(vertices=v, faces=f, face_mat=fm, materials=m, normals=nr)


def ref_ply(L, path):
    n = np.zeros(3, np.uint32)
    assert L.gvtref_ply_load(path.encode(), p(n)) == 0, path
    v, f, c = np.zeros((n[0], 3), np.float32), np.zeros((n[1], 3), np.uint32), np.zeros((n[0] if n[2] else 0, 3), np.float32)
    L.gvtref_ply_get(p(v), p(f), p(c))
    return dict(vertices=v, faces=f, colors=c)


def digest(d):
    return {k: {"shape": list(a.shape), "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()} for k, a in d.items()}


def write_binary_ply(path, big):
    """The tri_ascii.ply content as a binary PLY (double z, ushort list count) -- the same numbers in another encoding."""
    e = ">" if big else "<"
    verts = [(0, 0, 0, 255, 0, 0), (1, 0, 0, 0, 255, 0), (1, 1, 0, 0, 0, 255), (0, 1, 0, 128, 128, 128), (0.5, 0.5, 1.5, 10, 20, 30)]
    faces = [(0, 1, 2), (0, 2, 3), (0, 1, 2, 3)]
    hdr = ("ply\nformat binary_%s_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty double z\n"
           "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 3\nproperty list ushort uint vertex_indices\nend_header\n") % ("big" if big else "little")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        for x, y, z, r, g, b in verts:
            f.write(struct.pack(e + "ffdBBB", x, y, z, r, g, b))
        for fc in faces:
            f.write(struct.pack(e + "H%dI" % len(fc), len(fc), *fc))


FIXTURE_OBJ = ["box.obj"]
FIXTURE_PLY = ["tri_ascii.ply"]
REF_OBJ = ["data/geom/bunny.obj", "data/geom/obj/sibenik/sibenik.obj", "data/geom/cornell/cornell_box.obj", "data/geom/cube.obj",
           "third-party/tinyobjloader-master/cornell_box_multimaterial.obj", "third-party/tinyobjloader-master/cube.obj",
           "third-party/tinyobjloader-master/no_material.obj"]
REF_PLY = ["data/geom/bunny/reconstruction/bun_zipper.ply", "data/geom/bunny/reconstruction/bun_zipper_res4.ply"]


def main():
    L = ref_lib()
    out = {"fixtures": {}, "reference_files": {}}
    fx = os.path.join(HERE, "readers")
    for name in FIXTURE_OBJ:
        out["fixtures"][name] = digest(ref_obj(L, os.path.join(fx, name)))
    for name in FIXTURE_PLY:
        out["fixtures"][name] = digest(ref_ply(L, os.path.join(fx, name)))
    for name in REF_OBJ:
        if os.path.exists(os.path.join(REF, name)):
            out["reference_files"][name] = digest(ref_obj(L, os.path.join(REF, name)))
    for name in REF_PLY:
        if os.path.exists(os.path.join(REF, name)):
            out["reference_files"][name] = digest(ref_ply(L, os.path.join(REF, name)))
    json.dump(out, open(os.path.join(fx, "readers_golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(out["fixtures"]), "fixture and", len(out["reference_files"]), "reference-file digests")


if __name__ == "__main__":
    sys.exit(main())
