"""Convert the reference's scene files into compact binary assets (run once in the build container;
/root/reference does not exist on the GPU box, so the outputs are committed).

  data/geom/bunny.obj                              -> scenes/assets/bunny_obj.npz   (golden bunny.ppm scene)
  data/geom/bunny/reconstruction/bun_zipper.ply    -> scenes/assets/bunny_70k.npz   (BASELINE config 1)
  data/geom/obj/sibenik/sibenik.{obj,mtl}          -> scenes/assets/sibenik.npz     (BASELINE configs 2, 5)
  Test/CTESTtest/data/{simple,bunny}.ppm           -> tests/golden/*.ppm.gz         (reference golden images)

Faces are stored 0-based.  OBJ polygons are fan-triangulated (what tinyobjloader does for ObjReader.cpp).
"""
import gzip
import os
import shutil
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "assets")


def load_obj(path):
    verts, faces, fmat, names = [], [], [], []
    cur = -1
    for line in open(path, "r", errors="replace"):
        if line.startswith("v "):
            verts.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
            idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
            for k in range(1, len(idx) - 1):
                faces.append([idx[0], idx[k], idx[k + 1]])
                fmat.append(cur)
        elif line.startswith("usemtl"):
            name = line.split()[1]
            if name not in names:
                names.append(name)
            cur = names.index(name)
    return np.array(verts, np.float32), np.array(faces, np.uint32), np.array(fmat, np.int32), names


def load_mtl(path):
    mats, cur = {}, None
    for line in open(path, "r", errors="replace"):
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "newmtl":
            cur = tok[1]
            mats[cur] = {"Kd": [0.5, 0.5, 0.5], "Ks": [0.5, 0.5, 0.5], "Ns": 1.0}
        elif cur and tok[0] in ("Kd", "Ks"):
            mats[cur][tok[0]] = [float(x) for x in tok[1:4]]
        elif cur and tok[0] == "Ns":
            mats[cur]["Ns"] = float(tok[1])
    return mats


def load_ply_ascii(path):
    f = open(path, "r")
    nv = nf = 0
    while True:
        line = f.readline().strip()
        if line.startswith("element vertex"):
            nv = int(line.split()[-1])
        elif line.startswith("element face"):
            nf = int(line.split()[-1])
        elif line == "end_header":
            break
    verts = np.array([[float(x) for x in f.readline().split()[:3]] for _ in range(nv)], np.float32)
    faces = np.array([[int(x) for x in f.readline().split()[1:4]] for _ in range(nf)], np.uint32)
    return verts, faces


def main():
    os.makedirs(OUT, exist_ok=True)
    v, f, _, _ = load_obj(f"{REF}/data/geom/bunny.obj")
    np.savez_compressed(f"{OUT}/bunny_obj.npz", verts=v, faces=f)
    print("bunny_obj", v.shape, f.shape)

    v, f = load_ply_ascii(f"{REF}/data/geom/bunny/reconstruction/bun_zipper.ply")
    np.savez_compressed(f"{OUT}/bunny_70k.npz", verts=v, faces=f)
    print("bunny_70k", v.shape, f.shape)

    v, f, fm, names = load_obj(f"{REF}/data/geom/obj/sibenik/sibenik.obj")
    mtl = load_mtl(f"{REF}/data/geom/obj/sibenik/sibenik.mtl")
    mats = np.zeros((len(names), 8), np.float32)  # type, kd[3], ks[3], alpha
    for i, n in enumerate(names):
        m = mtl.get(n, {"Kd": [0.5] * 3, "Ks": [0.5] * 3, "Ns": 1.0})
        mats[i] = [0.0] + m["Kd"] + m["Ks"] + [m["Ns"]]
    np.savez_compressed(f"{OUT}/sibenik.npz", verts=v, faces=f, face_mat=fm, mats=mats)
    print("sibenik", v.shape, f.shape, len(names), "materials")

    gold = os.path.join(HERE, "..", "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    for name in ("simple", "bunny"):
        with open(f"{REF}/Test/CTESTtest/data/{name}.ppm", "rb") as src, gzip.open(f"{gold}/{name}.ppm.gz", "wb", 9) as dst:
            shutil.copyfileobj(src, dst)
    print("goldens copied")


if __name__ == "__main__":
    sys.exit(main())
