"""Scene ingestion: host/readers.cpp against the REAL third-party readers the reference links.

  * golden digests (tests/golden/readers/readers_golden.json, written by tests/golden/make_reader_golden.py from
    tinyobjloader / ply.c compiled out of /root/reference): committed fixtures everywhere, the reference's own scene files
    where /root/reference exists;
  * live, array by array, against oracle/_ref/libgvt_ref_readers.so when it is built;
  * the committed scene assets (scenes/assets/*.npz) describe the same triangles as the reference's files;
  * the OBJ number grammar is tinyobjloader's digit accumulation, not strtod;
  * ConfigFileLoader -> api:: database."""
import hashlib
import json
import math
import os
import sys

import numpy as np
import pytest

import gravit_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FX = os.path.join(ROOT, "tests", "golden", "readers")
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_reader_golden as mk  # noqa: E402

GOLD = json.load(open(os.path.join(FX, "readers_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def normalised(n):
    n = n.astype(np.float32)
    d = (n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2]).astype(np.float32)
    inv = (np.float32(1.0) / np.sqrt(d)).astype(np.float32)
    return (n * inv[:, None]).astype(np.float32)


def check_obj_digest(ours, gold):
    assert sha(ours["vertices"]) == gold["vertices"]["sha256"] and list(ours["vertices"].shape) == gold["vertices"]["shape"]
    assert sha(ours["faces"]) == gold["faces"]["sha256"]
    assert sha(ours["materials"][:, :9]) == gold["materials"]["sha256"]
    if gold["materials"]["shape"][0]:
        assert sha(ours["face_mat"]) == gold["face_mat"]["sha256"]
    assert list(ours["normals"].shape) == gold["normals"]["shape"]


@pytest.mark.parametrize("name", sorted(GOLD["fixtures"]))
def test_fixture_matches_reference_reader_digest(built, name):
    path = os.path.join(FX, name)
    if name.endswith(".obj"):
        check_obj_digest(g.read_obj(path), GOLD["fixtures"][name])
    else:
        o, gold = g.read_ply(path), GOLD["fixtures"][name]
        for k in ("vertices", "faces", "colors"):
            assert sha(o[k]) == gold[k]["sha256"], k


@pytest.mark.parametrize("name", sorted(GOLD["reference_files"]))
def test_reference_scene_file_matches_reference_reader_digest(built, name):
    path = os.path.join(REF, name)
    if not os.path.exists(path):
        pytest.skip("/root/reference is not on this host")
    if name.endswith(".obj"):
        check_obj_digest(g.read_obj(path), GOLD["reference_files"][name])
    else:
        o, gold = g.read_ply(path), GOLD["reference_files"][name]
        for k in ("vertices", "faces", "colors"):
            assert sha(o[k]) == gold[k]["sha256"], k


def test_live_against_compiled_reference_readers(built, tmp_path):
    so = os.path.join(ROOT, "oracle", "_ref", "libgvt_ref_readers.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref is not built on this host")
    L = mk.ref_lib()
    ours, ref = g.read_obj(os.path.join(FX, "box.obj")), mk.ref_obj(L, os.path.join(FX, "box.obj"))
    assert ours["vertices"].shape == (27, 3) or ours["vertices"].shape == ref["vertices"].shape
    for k in ("vertices", "faces", "face_mat"):
        assert np.array_equal(ours[k].view(np.uint32), ref[k].view(np.uint32)), k
    assert np.array_equal(ours["materials"][:, :9].view(np.uint32), ref["materials"].view(np.uint32))
    assert np.array_equal(ours["normals"].view(np.uint32), normalised(ref["normals"]).view(np.uint32))  # ObjReader.cpp:185
    # bounds = Mesh::computeBoundingBox over the flattened vertices
    assert np.array_equal(ours["bounds"][:3], ref["vertices"].min(0)) and np.array_equal(ours["bounds"][3:], ref["vertices"].max(0))
    # binary PLY through ply.c and through ours.  Little endian only: the vendored ply.c freads native byte order for
    # both binary formats, so it misreads big-endian files on x86; ours decodes them (test_binary_ply_equals_ascii)
    for big in (False,):
        path = str(tmp_path / f"b{int(big)}.ply")
        mk.write_binary_ply(path, big)
        o, r = g.read_ply(path), mk.ref_ply(L, path)
        for k in ("vertices", "faces", "colors"):
            assert np.array_equal(o[k].view(np.uint32), r[k].view(np.uint32)), (big, k)


def test_binary_ply_equals_ascii(built, tmp_path):
    a = g.read_ply(os.path.join(FX, "tri_ascii.ply"))
    assert a["faces"].tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]  # the first three indices of every list (PlyReader.cpp:165-167)
    assert a["colors"][0].tolist() == [1.0, 0.0, 0.0] and a["colors"][4].tolist() == [np.float32(10) / np.float32(255), np.float32(20) / np.float32(255), np.float32(30) / np.float32(255)]
    for big in (False, True):
        path = str(tmp_path / f"b{int(big)}.ply")
        mk.write_binary_ply(path, big)
        b = g.read_ply(path)
        for k in ("vertices", "faces", "colors", "bounds"):
            assert np.array_equal(a[k], b[k]), (big, k)


def tinyobj_number(text):
    """Independent restatement of tiny_obj_loader.cc:127-264 in Python doubles."""
    s, i, n = text, 0, len(text)
    sign = 1
    if s[i] in "+-":
        sign = -1 if s[i] == "-" else 1
        i += 1
    mant = 0.0
    while i < n and s[i].isdigit():
        mant = mant * 10 + int(s[i])
        i += 1
    if i < n and s[i] == ".":
        i += 1
        read = 1
        while i < n and s[i].isdigit():
            mant += int(s[i]) * math.pow(10.0, -read)
            read += 1
            i += 1
    expo = 0
    if i < n and s[i] in "eE":
        i += 1
        es = 1
        if s[i] in "+-":
            es = -1 if s[i] == "-" else 1
            i += 1
        while i < n and s[i].isdigit():
            expo = expo * 10 + int(s[i])
            i += 1
        expo *= es
    return np.float32(sign * math.ldexp(mant * math.pow(5.0, expo), expo))


def test_obj_number_grammar_is_tinyobj_not_strtod(built):
    rng = np.random.default_rng(7)
    texts = ["0", "-0", "+3.1417e+2", "-0.0E-3", "1.0324", "-1.41", "11e2", ".5", "abc", "1e", "5.0e-1", "0.50000001", "-13.9444", "1e-20", "3.4e38"]
    texts += ["%.*f" % (int(rng.integers(1, 12)), x) for x in rng.normal(0, 50, 300)]
    texts += ["%.*e" % (int(rng.integers(1, 9)), x) for x in rng.normal(0, 1e-3, 300)]
    # decimal strings just above the midpoint of two floats: a correctly rounded parser rounds up, digit accumulation may not
    mids = [repr(float(np.float32(x)) + float(np.spacing(np.float32(x))) / 2) for x in rng.uniform(0.1, 100, 400)]
    texts += [m[:-1] + str(min(9, int(m[-1]) + 1)) for m in mids]
    differs = 0
    for t in texts:
        ours = g.parse_obj_float(t)
        if t in ("abc", ".5", "1e"):  # not numbers in this grammar (no leading digit / empty exponent): 0
            assert ours == 0.0, t
            continue
        want = tinyobj_number(t)
        assert ours.view(np.uint32) == want.view(np.uint32), (t, ours, want)
        differs += int(np.float32(float(t)) != want)
    print("strings where strtod and the reference's parser disagree:", differs)


def soup(v, f):
    return v[f.astype(np.int64)].reshape(len(f), 9)


@pytest.mark.parametrize("asset,path", [("bunny_obj", "data/geom/bunny.obj"), ("sibenik", "data/geom/obj/sibenik/sibenik.obj")])
def test_committed_assets_describe_the_same_triangles(built, asset, path):
    """scenes/assets/*.npz were converted with a Python loader (scenes/make_assets.py); the C++ reader must give the same
    triangle soup in the same face order (vertex numbering differs: tinyobj re-indexes per shape)."""
    if not os.path.exists(os.path.join(REF, path)):
        pytest.skip("/root/reference is not on this host")
    a = np.load(os.path.join(ROOT, "scenes", "assets", asset + ".npz"))
    o = g.read_obj(os.path.join(REF, path))
    assert len(o["faces"]) == len(a["faces"])
    assert np.array_equal(soup(o["vertices"], o["faces"]), soup(a["verts"], a["faces"]))


def test_bunny_ply_asset(built):
    p = os.path.join(REF, "data/geom/bunny/reconstruction/bun_zipper.ply")
    if not os.path.exists(p):
        pytest.skip("/root/reference is not on this host")
    a = np.load(os.path.join(ROOT, "scenes", "assets", "bunny_70k.npz"))
    o = g.read_ply(p)
    assert np.array_equal(o["vertices"], a["verts"]) and np.array_equal(o["faces"], a["faces"]) and len(o["colors"]) == 0


def test_reader_errors(built, tmp_path):
    with pytest.raises(g.EngineError, match="Cannot open"):
        g.read_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(g.EngineError, match="out of range"):
        g.read_obj(str(bad))
    nomtl = tmp_path / "nomtl.obj"  # a missing library still yields one all-zero material (tiny_obj_loader.cc:700-703)
    nomtl.write_text("mtllib nothere.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl x\nf 1 2 3\n")
    with pytest.raises(g.EngineError, match="no mtllib defines"):
        g.read_obj(str(nomtl))
    with pytest.raises(g.EngineError, match="not a directory|does not exist"):
        g.load_ply_dir(str(tmp_path / "nodir"))
    with pytest.raises(g.EngineError, match="no .ply files"):
        g.load_ply_dir(str(tmp_path))


def test_conf_loader_builds_the_database(built):
    A = g.api()
    g.acheck(A.gvtapi_reset())
    info = g.load_conf(os.path.join(FX, "scene.conf"))
    assert info == dict(scheduler_type=0, domain_type=3, accel_type=1, width=96, height=64, n_instances=2, n_lights=1)
    m, minv, normi, box = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(9, np.float32), np.zeros(6, np.float32)
    g.acheck(A.gvtapi_instance_get(0, g._fp(m), g._fp(minv), g._fp(normi), g._fp(box)))
    assert m.reshape(4, 4).T.tolist() == [[1, 0, 0, np.float32(-0.8)], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    g.acheck(A.gvtapi_instance_get(1, g._fp(m), g._fp(minv), g._fp(normi), g._fp(box)))
    # translate then scale; the rotation triple has no effect (ConfigFileLoader.cpp:184-191)
    assert m.reshape(4, 4).T.tolist() == [[0.5, 0, 0, np.float32(0.8)], [0, 1.5, 0, 0], [0, 0, 0.5, 0], [0, 0, 0, 1]]
    # both G lines share one mesh (one adapter / BLAS per distinct file): faces as the reader produced them
    nf = np.zeros(1, np.uint32)
    g.acheck(A.gvtapi_mesh_get(b"Mesh0", g._fp(nf), None, None, None))
    assert nf[0] == len(g.read_obj(os.path.join(FX, "box.obj"))["faces"])
    assert A.gvtapi_mesh_get(b"Mesh1", g._fp(nf), None, None, None) != 0
    g.acheck(A.gvtapi_reset())


def test_ply_dir_loader(built, tmp_path):
    import shutil

    A = g.api()
    g.acheck(A.gvtapi_reset())
    shutil.copy(os.path.join(FX, "tri_ascii.ply"), tmp_path / "b.ply")
    mk.write_binary_ply(str(tmp_path / "a.ply"), False)
    assert g.load_ply_dir(str(tmp_path)) == 2
    nf = np.zeros(1, np.uint32)
    for name in (b"Mesh0", b"Mesh1"):
        g.acheck(A.gvtapi_mesh_get(name, g._fp(nf), None, None, None))
        assert nf[0] == 3
    box = np.zeros(6, np.float32)
    g.acheck(A.gvtapi_instance_get(1, None, None, None, g._fp(box)))
    assert box.tolist() == [0, 0, 0, 1, 1, 1.5]
    g.acheck(A.gvtapi_reset())
