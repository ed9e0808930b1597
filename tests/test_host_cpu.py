"""CPU-side checks: the C ABI library loads and exports every symbol include/gvtb200.h declares, fails loudly without
a device, and the C++ host layer (api:: mirror) reproduces the oracle's mesh / instance bookkeeping."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gravit_b200 as g
from oracle import pyoracle as po
from scenes import library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "gvtb200.h")).read()
    declared = sorted(set(re.findall(r"GVTB200_API\s+[\w\s\*]+?\b(gvtb200_\w+)\s*\(", header)))
    assert declared == sorted(g.C_ABI_SYMBOLS)
    L = g.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback(built):
    """Without a usable device every compute entry point refuses to work."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = g.lib()
    assert L.gvtb200_init(0) == -1  # GVTB200_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.gvtb200_last_error() or b"no CUDA device" in L.gvtb200_last_error()
    h = C.c_void_p()
    v = np.zeros((3, 3), np.float32)
    f = np.array([[0, 1, 2]], np.uint32)
    assert L.gvtb200_mesh_create(g._fp(v), 3, g._fp(f), 1, None, None, None, 0, None, None, C.byref(h)) == -1
    with pytest.raises(g.EngineError):
        s = library.apply(library.simple_scene(32, 32), g.GvtScene())
        s.render()


def test_oracle_is_not_linked_into_the_product(built):
    import subprocess

    for lib in (g.LIB_PATH, g.API_PATH):
        out = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
        assert "gvto_" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gravit_b200")):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "pyoracle" not in src and "libgvt_oracle" not in src, fn


@pytest.mark.parametrize("scene", ["simple", "bunny70k"])
def test_host_mesh_and_instances_match_oracle(built, scene):
    d = library.simple_scene(64, 64) if scene == "simple" else library.bunny_scene(64, 64)
    o = library.apply(d, po.OracleScene())
    s = g.GvtScene()
    ids = [s.add_mesh(m.verts, m.tris, one_based=m.one_based, kd=m.kd) for m in d.meshes]
    for mesh, m16 in d.instances:
        s.add_instance(ids[mesh], m16)
    for k, m in enumerate(d.meshes):
        nv = len(np.asarray(m.verts).reshape(-1, 3))
        fo, no, bo = o.mesh_get(k, nv)
        fh, nh, bh = s.mesh_get(k, nv)
        assert np.array_equal(fo.astype(np.uint32), fh)  # 1-based -> 0-based, degenerate faces dropped
        assert np.array_equal(no.view(np.uint32), nh.view(np.uint32))  # Mesh::generateNormals, bit for bit
        assert np.array_equal(bo, bh)
    for i in range(len(d.instances)):
        for a, b in zip(o.instance_get(i), s.instance_get(i)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))  # GLM inverse / normi / two-corner box


def test_add_face_bounds_and_unknown_names(built):
    s = g.GvtScene()
    with pytest.raises(g.EngineError, match="outside bounds"):
        s.add_mesh(np.zeros((3, 3), np.float32), np.array([[1, 2, 4]], np.uint32))  # Mesh.cpp:103-105
    A = g.api()
    assert A.gvtapi_add_instance(b"i0", b"nope", g._fp(library.identity())) == -1  # context.h:89 "not found"
    assert A.gvtapi_modify_camera(b"nope", None, None, None, 0.0, 1, 1, 0.0) == 0  # api.cpp:364-366: silently returns
    assert A.gvtapi_create_mesh(b"dup") == 0 and A.gvtapi_create_mesh(b"dup") == -1  # unique names


def test_degenerate_faces_are_dropped(built):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 0, 0]], np.float32)  # vertex 3 duplicates vertex 1
    t = np.array([[1, 2, 3], [1, 2, 4], [2, 3, 4]], np.uint32)
    s = g.GvtScene()
    s.add_mesh(v, t)
    faces, _, _ = s.mesh_get(0, 4)
    assert faces.tolist() == [[0, 1, 2]]  # faces 2 and 3 have two identical positions (Mesh.cpp:107-109)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present on this box")
def test_reference_side_adapter_shim_compiles_against_gravit_headers(tmp_path):
    """B200MeshAdapter.h is written against GraviT's own Adapter.h / Mesh.h / Light.h."""
    import subprocess

    src = tmp_path / "shim_check.cpp"
    src.write_text('#include "B200MeshAdapter.h"\nint main() { return 0; }\n')
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-w", "-I/root/reference/src", "-I/root/reference/third-party/glm",
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "gravit_b200", "host", "gvt_shim"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:2000]
