"""pygvt-compatible binding and the adapter-type enumerators (N1: Embree vs EmbreeStream).

The function list below is the `def` list of the reference's pygvt/src/gvt/gvt.pyx:49-140 (committed here as names
because /root/reference does not travel to the GPU box); the reference has no tests for pygvt, so its example scripts
(pygvt/*.py) and SimpleApp.cpp's scene are the model."""
import numpy as np
import pytest

from oracle import pyoracle as po
from scenes import library

from .scene_util import golden_ppm, oracle_scene

GVT_PYX_DEFS = (
    "gvtInit gvtsync createMesh addVolumeTransferFunctions addVolumeSamples addAmrSubgrid addMeshVertices addMeshTriangles "
    "addMeshFaceNormals addMeshVertexNormals finishMesh addMeshMaterialLambert addMeshMaterialSpecular addInstance createVolume "
    "addPointLight addAreaLight modifyLight modifyLight2 addCamera modifyCamera addFilm modifyFilm addRenderer render writeimage"
).split()


def test_pygvt_surface_and_argument_checks():
    import gravit_b200.pygvt as gvt

    for name in GVT_PYX_DEFS:
        assert callable(getattr(gvt, name)), name
    # Types.h:33-46,50-62 enumerator values
    assert (gvt.Surface, gvt.Manta, gvt.Optix, gvt.Embree, gvt.Ospray, gvt.Pvol, gvt.EmbreeStream, gvt.Heterogeneous) == tuple(range(8))
    assert (gvt.Image, gvt.Domain, gvt.AsyncImage, gvt.AsyncDomain) == (0, 1, 2, 3)
    gvt.gvtReset()
    gvt.createMesh("m")
    with pytest.raises(RuntimeError):  # duplicate names are refused by the database (core/cntx/context.h:241)
        gvt.createMesh("m")
    with pytest.raises(RuntimeError):  # unknown mesh
        gvt.addMeshVertices("nope", 1, np.zeros(3, np.float32))
    with pytest.raises(ValueError):  # Cython typed buffers: float64 is a dtype mismatch
        gvt.addMeshVertices("m", 1, np.zeros(3, np.float64))
    with pytest.raises(ValueError):  # mode="c"
        gvt.addMeshVertices("m", 1, np.zeros(6, np.float32)[::2])
    with pytest.raises(TypeError):
        gvt.addMeshVertices("m", 1, None)
    with pytest.raises(ValueError):  # fewer values than `size` says would be read
        gvt.addMeshVertices("m", 2, np.zeros(3, np.float32))
    with pytest.raises(NotImplementedError):
        gvt.createVolume("v")
    gvt.addMeshVertices("m", 3, np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32))
    gvt.addMeshTriangles("m", 1, np.array([1, 2, 3], np.uint32))
    gvt.finishMesh("m")
    gvt.modifyLight("nolight", np.zeros(3, np.float32), np.zeros(3, np.float32))  # silent on unknown names (api.cpp:364-366)
    gvt.gvtReset()


def _build_simple_with_pygvt(gvt, adapter, w, h):
    """SimpleApp.cpp:82-233 written the way pygvt/example scripts write scenes."""
    d = library.simple_scene(w, h)
    gvt.gvtReset()
    gvt.gvtInit()
    for i, m in enumerate(d.meshes):
        name = f"mesh{i}"
        gvt.createMesh(name)
        v = np.ascontiguousarray(m.verts, np.float32).reshape(-1)
        f = np.ascontiguousarray(m.tris, np.uint32).reshape(-1)
        gvt.addMeshVertices(name, len(v) // 3, v)
        gvt.addMeshTriangles(name, len(f) // 3, f)
        gvt.addMeshMaterialLambert(name, gvt.LAMBERT, np.asarray(m.kd, np.float32), 1.0)
        gvt.finishMesh(name, True)
    for k, (mesh, m16) in enumerate(d.instances):
        gvt.addInstance(f"inst{k}", f"mesh{mesh}", np.ascontiguousarray(m16, np.float32))
    for k, (pos, col) in enumerate(d.lights):
        gvt.addPointLight(f"light{k}", np.asarray(pos, np.float32), np.asarray(col, np.float32))
    gvt.addCamera("Camera", np.asarray(d.eye, np.float32), np.asarray(d.focus, np.float32), np.asarray(d.up, np.float32), d.fov, d.depth, d.samples, d.jitter)
    gvt.addFilm("Film", w, h, "simple")
    gvt.addRenderer("Renderer", adapter, gvt.Image, "Camera", "Film", False)
    gvt.gvtsync()
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("adapter,smooth", [("Embree", 0), ("EmbreeStream", 1), ("B200", 0)])
def test_pygvt_simple_scene_adapter_types(gpu, adapter, smooth, tmp_path):
    """adapter::Embree is HEAD's FLAT_SHADING build (EmbreeMeshAdapter.cpp:75); adapter::EmbreeStream has no such define
    (EmbreeStreamMeshAdapter.cpp:931-946) and therefore reproduces the reference's golden image, which predates the
    define.  Both must equal the oracle bit for bit."""
    import gravit_b200.pygvt as gvt

    d = _build_simple_with_pygvt(gvt, getattr(gvt, adapter), 512, 512)
    gvt.render("Renderer")
    fb = gvt.framebuffer()
    ref, _ = oracle_scene(d, flags=smooth).render(threads=8)
    assert np.array_equal(fb, ref)
    out = str(tmp_path / "simple")
    gvt.writeimage("Renderer", out)
    w, h, px = po.read_ppm(open(out + ".ppm", "rb").read())
    assert (w, h) == (512, 512)
    assert po.image_diff(px, po.fb_to_ppm_bytes(ref)) == 0
    if smooth:
        _, _, gold = golden_ppm("simple")
        assert po.image_diff(px, gold) < 300 and po.psnr(px, gold) >= 50.0  # CMakeLists.txt:666
    # modifyCamera (4-argument form) + modifyFilm + re-render: the same database, a new view
    gvt.modifyCamera("Camera", np.array([4, 0.5, 0], np.float32), np.asarray(d.focus, np.float32), np.asarray(d.up, np.float32), d.fov)
    gvt.modifyFilm("Film", 256, 256, "simple2")
    gvt.render("Renderer")
    fb2 = gvt.framebuffer()
    d.eye, d.width, d.height = (4, 0.5, 0), 256, 256
    ref2, _ = oracle_scene(d, flags=smooth).render(threads=8)
    assert np.array_equal(fb2, ref2)
    gvt.gvtReset()


def test_pygvt_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    import gravit_b200.pygvt as gvt

    gvt.gvtReset()
    with pytest.raises(RuntimeError):
        gvt.gvtInit()
