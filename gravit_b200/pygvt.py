"""pygvt-compatible Python binding: the function names and argument lists of the reference's Cython module
`gvt` (/root/reference/pygvt/src/gvt/gvt.pyx:49-140), bound with ctypes to the C entry points of the C++ `api::` mirror
(gravit_b200/host/api.cpp -> libgvtb200_api.so) instead of GraviT's libgvtRender.  A reference script that does

    import gvt
    gvt.gvtInit(); gvt.createMesh("m"); gvt.addMeshVertices("m", n, verts) ... gvt.render("r"); gvt.writeimage("r", "out")

runs unchanged with `import gravit_b200.pygvt as gvt`.  Arrays are flat C-contiguous numpy arrays of the dtype the .pyx
declares (float32 / uint32); like the Cython typed-buffer arguments, a wrong dtype or a non-contiguous array raises
ValueError and `None` raises TypeError.  The volume entry points (createVolume, addVolumeSamples, addAmrSubgrid,
addVolumeTransferFunctions: gvt.pyx:73-78,104-105) raise NotImplementedError: volume rendering is outside the hot path
(DESIGN.md §6).  Engine errors surface as RuntimeError (the reference lets C++ exceptions terminate the interpreter).

Additions (not in gvt.pyx): `framebuffer(name)` returns the RGBA float image of the last render, `setEngineFlags`.
"""
import ctypes as C

import numpy as np

from . import EngineError, api

# gvt::render::adapter::RenderType / scheduler::ScheduleType (render/Types.h:33-62) for scripts that name them
Surface, Manta, Optix, Embree, Ospray, Pvol, EmbreeStream, Heterogeneous, B200 = range(9)
Image, Domain, AsyncImage, AsyncDomain = range(4)
LAMBERT, PHONG, BLINN = 0, 1, 2  # data/primitives/Material.h:44-57


def _arr(a, dtype, name, min_len=0):
    if a is None:
        raise TypeError(f"Argument '{name}' must not be None")
    if not isinstance(a, np.ndarray) or a.ndim != 1 or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"]:
        raise ValueError(f"Buffer dtype mismatch or not a 1-d C-contiguous array: '{name}' must be ndarray[{np.dtype(dtype).name}, ndim=1, mode='c']")
    if a.size < min_len:
        raise ValueError(f"'{name}' holds {a.size} values, {min_len} are read")
    return a.ctypes.data_as(C.c_void_p)


def _ck(rc):
    if rc != 0:
        raise RuntimeError(api().gvtapi_last_error().decode())


def _s(x):
    if not isinstance(x, str):
        raise TypeError("expected str")
    return x.encode()


def gvtInit():
    _ck(api().gvtapi_init())


def gvtsync():
    _ck(api().gvtapi_gvtsync())


def createMesh(name):
    _ck(api().gvtapi_create_mesh(_s(name)))


def addMeshVertices(name, size, vertices):
    _ck(api().gvtapi_add_mesh_vertices(_s(name), int(size), _arr(vertices, np.float32, "vertices", 3 * int(size))))


def addMeshTriangles(name, n, triangles):
    """1-based vertex indices (api.cpp:173-180 -> Mesh::addFace)."""
    _ck(api().gvtapi_add_mesh_triangles(_s(name), int(n), _arr(triangles, np.uint32, "triangles", 3 * int(n))))


def addMeshFaceNormals(name, n, normals):
    _ck(api().gvtapi_add_mesh_face_normals(_s(name), int(n), _arr(normals, np.float32, "normals", 3 * int(n))))


def addMeshVertexNormals(name, n, normals):
    _ck(api().gvtapi_add_mesh_vertex_normals(_s(name), int(n), _arr(normals, np.float32, "normals", 3 * int(n))))


def finishMesh(name, compute_normal=True):
    _ck(api().gvtapi_finish_mesh(_s(name), 1 if compute_normal else 0))


def addMeshMaterialLambert(name, mattype, kd, alpha):
    _ck(api().gvtapi_add_mesh_material(_s(name), int(mattype), _arr(kd, np.float32, "kd", 3), None, float(alpha)))


def addMeshMaterialSpecular(name, mattype, kd, ks, alpha):
    _ck(api().gvtapi_add_mesh_material(_s(name), int(mattype), _arr(kd, np.float32, "kd", 3), _arr(ks, np.float32, "ks", 3), float(alpha)))


def addInstance(s, name, m):
    """m: 16 floats, column-major (glm::make_mat4, api.cpp:292-322)."""
    _ck(api().gvtapi_add_instance(_s(s), _s(name), _arr(m, np.float32, "m", 16)))


def addPointLight(name, pos, color):
    _ck(api().gvtapi_add_point_light(_s(name), _arr(pos, np.float32, "pos", 3), _arr(color, np.float32, "color", 3)))


def addAreaLight(name, pos, color, n, w, h):
    _ck(api().gvtapi_add_area_light(_s(name), _arr(pos, np.float32, "pos", 3), _arr(color, np.float32, "color", 3), _arr(n, np.float32, "n", 3), float(w), float(h)))


def modifyLight(name, pos, color):
    _ck(api().gvtapi_modify_light(_s(name), _arr(pos, np.float32, "pos", 3), _arr(color, np.float32, "color", 3)))


def modifyLight2(name, pos, color, n, w, h):
    _ck(api().gvtapi_modify_area_light(_s(name), _arr(pos, np.float32, "pos", 3), _arr(color, np.float32, "color", 3), _arr(n, np.float32, "n", 3), float(w), float(h)))


def addCamera(name, pos, focus, up, fov, depth, samples, jitter):
    _ck(api().gvtapi_add_camera(_s(name), _arr(pos, np.float32, "pos", 3), _arr(focus, np.float32, "focus", 3), _arr(up, np.float32, "up", 3),
                                float(fov), int(depth), int(samples), float(jitter)))


def modifyCamera(name, pos, focus, up, fov):
    _ck(api().gvtapi_modify_camera_pose(_s(name), _arr(pos, np.float32, "pos", 3), _arr(focus, np.float32, "focus", 3), _arr(up, np.float32, "up", 3), float(fov)))


def addFilm(name, w, h, path):
    _ck(api().gvtapi_add_film(_s(name), int(w), int(h), _s(path)))


def modifyFilm(name, w, h, path):
    _ck(api().gvtapi_modify_film(_s(name), int(w), int(h), _s(path)))


def addRenderer(name, adapter, schedule, camera_name, film_name, volume):
    _ck(api().gvtapi_add_renderer_volume(_s(name), int(adapter), int(schedule), _s(camera_name), _s(film_name), 1 if volume else 0))


def render(name):
    _ck(api().gvtapi_render(_s(name)))


def writeimage(name, output):
    """Writes `output`.ppm (IceTComposite::write, composite/IceTComposite.cpp:119-158)."""
    _ck(api().gvtapi_writeimage(_s(name), _s(output)))


def _no_volume(*_a, **_k):
    raise NotImplementedError("volume rendering (createVolume/addVolumeSamples/addAmrSubgrid/addVolumeTransferFunctions) is outside this engine's hot path")


createVolume = addVolumeTransferFunctions = addVolumeSamples = addAmrSubgrid = _no_volume


# ---- additions ------------------------------------------------------------------------------------------------------
def framebuffer():
    """(H, W, 4) float32 image of the last render (gvtRenderer's myimage->colorbf(), composite/IceTComposite.h:106)."""
    w, h = C.c_int(0), C.c_int(0)
    _ck(api().gvtapi_framebuffer(None, 0, C.byref(w), C.byref(h)))
    fb = np.zeros((h.value, w.value, 4), np.float32)
    _ck(api().gvtapi_framebuffer(fb.ctypes.data_as(C.c_void_p), fb.size, None, None))
    return fb


def setEngineFlags(flags, seed=0):
    _ck(api().gvtapi_set_engine_flags(int(flags), int(seed)))


def gvtReset():
    _ck(api().gvtapi_reset())


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("C", "np", "api", "EngineError")]
