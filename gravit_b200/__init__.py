"""gravit_b200 -- B200-native render adapter for TACC/GraviT.

Python here is plumbing only: it loads the two in-tree shared libraries and exposes them to tests, bench.py and the
multi-GPU drivers.

  libgvtb200.so      C ABI (include/gvtb200.h) over the sm_100a CUDA kernels (gravit_b200/csrc)
  libgvtb200_api.so  C++ host layer mirroring gvt::render::api (gravit_b200/host/api.h) + its C entry points

There is no CPU fallback: if the CUDA library is missing, or no sm_100 device is visible, every compute call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GVTB200_LIB_PATH") or os.path.join(HERE, "libgvtb200.so")  # the override is a measurement hook (kernel variants)
API_PATH = os.path.join(HERE, "libgvtb200_api.so")

# gvt::render::actor::Ray (actor/Ray.h:68-95): 64-byte payload in an 80-byte record
RAY_DTYPE = np.dtype(
    [
        ("origin", "<f4", 3),
        ("t_min", "<f4"),
        ("direction", "<f4", 3),
        ("t_max", "<f4"),
        ("color", "<f4", 3),
        ("t", "<f4"),
        ("id", "<i4"),
        ("depth", "<i4"),
        ("w", "<f4"),
        ("type", "<i4"),
        ("pad", "u1", 16),
    ]
)
HIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
LIGHT_DTYPE = np.dtype(
    [("type", "<i4"), ("position", "<f4", 3), ("color", "<f4", 3), ("normal", "<f4", 3), ("width", "<f4"), ("height", "<f4")]
)
MATERIAL_DTYPE = np.dtype([("type", "<i4"), ("kd", "<f4", 3), ("ks", "<f4", 3), ("alpha", "<f4")])
COUNTER_FIELDS = (
    "camera_rays routed_initial primary_hits shadow_generated shadow_unoccluded secondary_generated fb_adds "
    "instance_visits nodes_visited tris_tested rays_exported stack_overflows queue_overflows"
).split()

PRIMARY, SHADOW, SECONDARY = 0, 1, 2
FLAG_SMOOTH, FLAG_ENTER_INSIDE, FLAG_SHADOW_TO_LIGHT = 1, 2, 8
SCHEDULE_IMAGE, SCHEDULE_DOMAIN = 0, 1
ADAPTER_B200 = 8  # gvt::render::adapter::B200 (gravit_b200/host/api.h)

# every symbol include/gvtb200.h declares (tests check that the library exports all of them)
C_ABI_SYMBOLS = (
    "gvtb200_init gvtb200_last_error gvtb200_current_device gvtb200_device_info gvtb200_set_tuning gvtb200_set_tail_split gvtb200_set_tail_prefetch gvtb200_set_peer_timeout_ms gvtb200_mesh_create gvtb200_mesh_destroy gvtb200_mesh_morton "
    "gvtb200_mesh_stats gvtb200_trace gvtb200_intersect gvtb200_occluded gvtb200_scene_create gvtb200_scene_destroy "
    "gvtb200_scene_add_instance gvtb200_scene_set_lights gvtb200_scene_set_camera gvtb200_scene_set_film "
    "gvtb200_scene_set_flags gvtb200_scene_set_queue_limit gvtb200_scene_commit gvtb200_scene_tlas_order gvtb200_camera_rays gvtb200_route "
    "gvtb200_shuffle gvtb200_framebuffer_clear gvtb200_framebuffer_read gvtb200_framebuffer_read_rows gvtb200_render_frame gvtb200_render_frame_async "
    "gvtb200_render_frame_tiles gvtb200_scene_set_stream gvtb200_launch_count gvtb200_scene_sync gvtb200_scene_counters gvtb200_scene_profile_counters gvtb200_scene_profile_launches gvtb200_scene_launch_records gvtb200_scene_device_framebuffer "
    "gvtb200_framebuffer_clamp gvtb200_scene_last_timing gvtb200_domain_set_owner gvtb200_domain_begin_frame "
    "gvtb200_domain_drain_local gvtb200_domain_export gvtb200_domain_import gvtb200_domain_pending "
    "gvtb200_p2p_export gvtb200_p2p_attach gvtb200_p2p_detach gvtb200_image_frame_p2p gvtb200_domain_round_p2p gvtb200_domain_composite_p2p "
    "gvtb200_domain_frame_persist gvtb200_domain_persist_config gvtb200_domain_persist_stats gvtb200_domain_persist_profile gvtb200_domain_persist_progress gvtb200_domain_set_servers "
    "gvtb200_scene_set_instance_mesh gvtb200_mesh_export gvtb200_mesh_import"
).split()
PEER_DESC_BYTES = 64 * 4 + 8 * 4 + 8 + 8 + 8 + 4 + 4  # sizeof(gvtb200_peer_desc)
MESH_DESC_BYTES = 6 * 64 + 6 * 8 + 6 * 8 + 8 + 32 + 16 + 8  # sizeof(gvtb200_mesh_desc)
IMPORT_COPY, IMPORT_MAP = 0, 1


class EngineError(RuntimeError):
    pass


def source_hash():
    """sha256 over the kernel sources (gravit_b200/csrc): ties a committed ncu capture to the code it was taken from."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(HERE, "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_lib = None
_api = None


def lib():
    """The C ABI library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.gvtb200_last_error.restype = C.c_char_p
        vp, sz, i32, u32, f = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_float
        L.gvtb200_init.argtypes = [i32]
        L.gvtb200_device_info.argtypes = [vp, vp, vp, vp]
        L.gvtb200_set_tuning.argtypes = [i32, i32, i32, i32]
        L.gvtb200_mesh_create.argtypes = [vp, u32, vp, u32, vp, vp, vp, u32, vp, vp, vp]
        L.gvtb200_mesh_destroy.argtypes = [vp]
        L.gvtb200_mesh_morton.argtypes = [vp, vp, vp]
        L.gvtb200_mesh_stats.argtypes = [vp, vp]
        L.gvtb200_trace.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp, vp, vp, vp, vp, i32, u32, u32]
        L.gvtb200_intersect.argtypes = [vp, vp, sz, vp]
        L.gvtb200_occluded.argtypes = [vp, vp, sz, vp]
        L.gvtb200_scene_create.argtypes = [vp]
        L.gvtb200_scene_destroy.argtypes = [vp]
        L.gvtb200_scene_add_instance.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.gvtb200_scene_set_lights.argtypes = [vp, vp, i32]
        L.gvtb200_scene_set_camera.argtypes = [vp, vp, vp, vp, f, i32, i32, f]
        L.gvtb200_scene_set_film.argtypes = [vp, i32, i32]
        L.gvtb200_scene_set_flags.argtypes = [vp, u32, u32]
        L.gvtb200_scene_commit.argtypes = [vp]
        L.gvtb200_scene_set_queue_limit.argtypes = [vp, sz]
        L.gvtb200_scene_tlas_order.argtypes = [vp, vp, vp]
        L.gvtb200_camera_rays.argtypes = [vp, vp, sz]
        L.gvtb200_route.argtypes = [vp, vp, sz, i32, vp, vp]
        L.gvtb200_shuffle.argtypes = [vp, vp, sz, i32, vp, vp, sz, vp]
        L.gvtb200_framebuffer_clear.argtypes = [vp]
        L.gvtb200_framebuffer_read.argtypes = [vp, vp, sz]
        L.gvtb200_framebuffer_clamp.argtypes = [vp]
        L.gvtb200_framebuffer_read_rows.argtypes = [vp, vp, i32, i32, i32, i32]
        L.gvtb200_render_frame.argtypes = [vp, i32, i32, vp, vp]
        L.gvtb200_render_frame_async.argtypes = [vp, i32, i32]
        L.gvtb200_scene_sync.argtypes = [vp]
        L.gvtb200_render_frame_tiles.argtypes = [vp, i32, i32, i32, i32, i32]
        L.gvtb200_scene_set_stream.argtypes = [vp, vp]
        L.gvtb200_launch_count.restype = C.c_ulonglong
        L.gvtb200_scene_counters.argtypes = [vp, vp]
        L.gvtb200_scene_profile_counters.argtypes = [vp, i32]
        L.gvtb200_scene_profile_launches.argtypes = [vp, i32]
        L.gvtb200_scene_launch_records.argtypes = [vp, i32, vp, vp, vp, vp]
        L.gvtb200_scene_device_framebuffer.argtypes = [vp, vp, vp]
        L.gvtb200_scene_last_timing.argtypes = [vp, vp, vp]
        L.gvtb200_domain_set_owner.argtypes = [vp, vp, i32, i32, i32]
        L.gvtb200_domain_begin_frame.argtypes = [vp]
        L.gvtb200_domain_drain_local.argtypes = [vp]
        L.gvtb200_domain_export.argtypes = [vp, vp, vp, vp]
        L.gvtb200_domain_import.argtypes = [vp, vp, sz]
        L.gvtb200_domain_pending.argtypes = [vp, vp]
        L.gvtb200_p2p_export.argtypes = [vp, i32, vp]
        L.gvtb200_p2p_attach.argtypes = [vp, i32, i32, vp]
        L.gvtb200_p2p_detach.argtypes = [vp]
        L.gvtb200_image_frame_p2p.argtypes = [vp, i32, i32, i32, i32, i32]
        L.gvtb200_domain_round_p2p.argtypes = [vp, i32, vp]
        L.gvtb200_domain_composite_p2p.argtypes = [vp, i32]
        L.gvtb200_set_peer_timeout_ms.argtypes = [i32]
        L.gvtb200_set_tail_prefetch.argtypes = [i32]
        L.gvtb200_set_tail_split.argtypes = [i32]
        L.gvtb200_domain_frame_persist.argtypes = [vp, i32, i32]
        L.gvtb200_domain_persist_config.argtypes = [vp, i32, i32]
        L.gvtb200_domain_persist_stats.argtypes = [vp, vp, vp, vp, i32]
        L.gvtb200_domain_set_servers.argtypes = [vp, vp, vp, i32]
        L.gvtb200_domain_persist_profile.argtypes = [vp, vp]
        L.gvtb200_domain_persist_progress.argtypes = [vp, vp, i32, vp]
        L.gvtb200_scene_set_instance_mesh.argtypes = [vp, i32, vp]
        L.gvtb200_mesh_export.argtypes = [vp, vp]
        L.gvtb200_mesh_import.argtypes = [vp, i32, vp]
        _lib = L
    return _lib


def api():
    """The C++ host layer (gvt::render::api mirror) through its C entry points."""
    global _api
    if _api is None:
        lib()
        if not os.path.exists(API_PATH):
            raise EngineError(f"{API_PATH} is missing: run __graft_entry__.build()")
        A = C.CDLL(API_PATH)
        A.gvtapi_last_error.restype = C.c_char_p
        vp, i32, u32, f, s = C.c_void_p, C.c_int, C.c_uint, C.c_float, C.c_char_p
        A.gvtapi_create_mesh.argtypes = [s]
        A.gvtapi_add_mesh_vertices.argtypes = [s, u32, vp]
        A.gvtapi_add_mesh_triangles.argtypes = [s, u32, vp]
        A.gvtapi_add_mesh_triangles0.argtypes = [s, u32, vp]
        A.gvtapi_add_mesh_vertex_normals.argtypes = [s, u32, vp]
        A.gvtapi_finish_mesh.argtypes = [s, i32]
        A.gvtapi_add_mesh_material.argtypes = [s, u32, vp, vp, f]
        A.gvtapi_add_mesh_materials.argtypes = [s, u32, vp, vp, vp, vp]
        A.gvtapi_add_mesh_vertex_color.argtypes = [s, u32, vp]
        A.gvtapi_add_instance.argtypes = [s, s, vp]
        A.gvtapi_add_point_light.argtypes = [s, vp, vp]
        A.gvtapi_add_area_light.argtypes = [s, vp, vp, vp, f, f]
        A.gvtapi_modify_light.argtypes = [s, vp, vp]
        A.gvtapi_add_camera.argtypes = [s, vp, vp, vp, f, i32, i32, f]
        A.gvtapi_modify_camera.argtypes = [s, vp, vp, vp, f, i32, i32, f]
        A.gvtapi_add_film.argtypes = [s, i32, i32, s]
        A.gvtapi_modify_film.argtypes = [s, i32, i32, s]
        A.gvtapi_add_renderer.argtypes = [s, i32, i32, s, s]
        A.gvtapi_modify_area_light.argtypes = [s, vp, vp, vp, f, f]
        A.gvtapi_add_mesh_face_normals.argtypes = [s, u32, vp]
        A.gvtapi_modify_renderer.argtypes = [s, i32, i32, s, s]
        A.gvtapi_add_renderer_volume.argtypes = [s, i32, i32, s, s, i32]
        A.gvtapi_modify_camera_pose.argtypes = [s, vp, vp, vp, f]
        A.gvtapi_render.argtypes = [s]
        A.gvtapi_set_gpus.argtypes = [i32]
        A.gvtapi_writeimage.argtypes = [s, s]
        A.gvtapi_set_engine_flags.argtypes = [u32, u32]
        A.gvtapi_framebuffer.argtypes = [vp, C.c_size_t, vp, vp]
        A.gvtapi_engine_scene.argtypes = [s]
        A.gvtapi_engine_scene.restype = vp
        A.gvtapi_instance_get.argtypes = [i32, vp, vp, vp, vp]
        A.gvtapi_mesh_get.argtypes = [s, vp, vp, vp, vp]
        # scene ingestion (host/readers.cpp)
        A.gvtapi_reader_last_error.restype = C.c_char_p
        A.gvtapi_obj_read.argtypes = [s, i32, vp, vp, vp, vp, vp, vp, vp]
        A.gvtapi_ply_read.argtypes = [s, vp, vp, vp, vp, vp]
        A.gvtapi_parse_float.argtypes = [s, vp]
        A.gvtapi_load_obj.argtypes = [s, i32, s]
        A.gvtapi_load_ply_dir.argtypes = [s, i32, vp]
        A.gvtapi_load_conf.argtypes = [s, vp]
        _api = A
    return _api


def check(rc):
    if rc < 0:
        raise EngineError(f"gvtb200 error {rc}: {lib().gvtb200_last_error().decode()}")
    return rc


def acheck(rc):
    if rc != 0:
        raise EngineError(api().gvtapi_last_error().decode())


def rcheck(rc):
    if rc != 0:
        raise EngineError(api().gvtapi_reader_last_error().decode())


def read_obj(path, material_type=0):
    """ObjReader (host/readers.cpp) without scene calls: dict of vertices, faces (0-based), face_mat, materials
    (ka, kd, ks, type), normals as stored in the file, bounds."""
    A = api()
    n = np.zeros(5, np.uint32)
    rcheck(A.gvtapi_obj_read(path.encode(), material_type, _fp(n), None, None, None, None, None, None))
    v, f = np.zeros((n[0], 3), np.float32), np.zeros((n[1], 3), np.uint32)
    fm, mats = np.zeros(n[1] if n[2] else 0, np.int32), np.zeros((n[2], 10), np.float32)
    nr, b = np.zeros((n[3], 3), np.float32), np.zeros(6, np.float32)
    rcheck(A.gvtapi_obj_read(path.encode(), material_type, _fp(n), _fp(v), _fp(f), _fp(fm), _fp(mats), _fp(nr), _fp(b)))
    return dict(vertices=v, faces=f, face_mat=fm, materials=mats, normals=nr, bounds=b)


def read_ply(path):
    """PlyReader::readFile: vertices, faces (0-based, first three indices of every list), colors in [0,1] or empty."""
    A = api()
    n = np.zeros(5, np.uint32)
    rcheck(A.gvtapi_ply_read(path.encode(), _fp(n), None, None, None, None))
    v, f, c, b = np.zeros((n[0], 3), np.float32), np.zeros((n[1], 3), np.uint32), np.zeros((n[4], 3), np.float32), np.zeros(6, np.float32)
    rcheck(A.gvtapi_ply_read(path.encode(), _fp(n), _fp(v), _fp(f), _fp(c), _fp(b)))
    return dict(vertices=v, faces=f, colors=c, bounds=b)


def parse_obj_float(text):
    out = np.zeros(1, np.float32)
    api().gvtapi_parse_float(text.encode(), _fp(out))
    return out[0]


def load_obj(path, meshname, material_type=0):
    """ObjReader(path).addToScene(meshname): the mesh goes into the api:: database."""
    rcheck(api().gvtapi_load_obj(path.encode(), material_type, meshname.encode()))


def load_ply_dir(path, add_instances=True):
    n = C.c_int(0)
    rcheck(api().gvtapi_load_ply_dir(path.encode(), int(add_instances), C.byref(n)))
    return n.value


def load_conf(path):
    """gvtapps::render::ConfigFileLoader: builds the whole scene (meshes, instances, lights, camera "Camera", film "Film",
    renderer "Renderer") in the api:: database."""
    info = np.zeros(7, np.int32)
    rcheck(api().gvtapi_load_conf(path.encode(), _fp(info)))
    return dict(zip(("scheduler_type", "domain_type", "accel_type", "width", "height", "n_instances", "n_lights"), (int(x) for x in info)))


_inited = None


def init(device=0):
    """gvtb200_init; raises EngineError when no sm_100 device is usable (no CPU fallback)."""
    global _inited
    if _inited != device:
        check(lib().gvtb200_init(device))
        _inited = device
    return device


def device_info():
    sm, hbm, maj, mnr = C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
    check(lib().gvtb200_device_info(C.byref(sm), C.byref(hbm), C.byref(maj), C.byref(mnr)))
    return {"sm_count": sm.value, "hbm_bytes": hbm.value, "cc": (maj.value, mnr.value)}


class Mesh:
    """gvtb200_mesh handle == one engine adapter object (reference: EmbreeMeshAdapter, one per mesh)."""

    def __init__(self, verts, faces0, normals=None, material=None, mats=None, face_mat=None, vcolors=None):
        init(_inited if _inited is not None else 0)
        v = f32(verts).reshape(-1, 3)
        f = np.ascontiguousarray(faces0, dtype=np.uint32).reshape(-1, 3)
        n = None if normals is None else f32(normals).reshape(-1, 3)
        mm = None
        if material is not None:
            mm = np.zeros(1, MATERIAL_DTYPE)
            mm[0] = material
        mt = None if mats is None else np.ascontiguousarray(mats, dtype=MATERIAL_DTYPE)
        fm = None if face_mat is None else np.ascontiguousarray(face_mat, dtype=np.int32)
        vc = None if vcolors is None else f32(vcolors).reshape(-1, 3)
        self.h = C.c_void_p()
        self.nv, self.nf = len(v), len(f)
        check(
            lib().gvtb200_mesh_create(
                _fp(v), len(v), _fp(f), len(f), _fp(n), _fp(mm), _fp(mt), 0 if mt is None else len(mt), _fp(fm), _fp(vc), C.byref(self.h)
            )
        )

    def close(self):
        if self.h:
            lib().gvtb200_mesh_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self):
        """gvtb200_mesh_desc as bytes: lets another rank import this mesh's BLAS (replica or NVLink mapping)."""
        buf = C.create_string_buffer(MESH_DESC_BYTES)
        check(lib().gvtb200_mesh_export(self.h, buf))
        return buf.raw

    @classmethod
    def from_desc(cls, desc, mode=IMPORT_COPY):
        """The BLAS another rank built: mode IMPORT_COPY = replica in this GPU's HBM, IMPORT_MAP = traverse the owner's arrays."""
        desc = bytes(desc)
        if len(desc) != MESH_DESC_BYTES:
            raise EngineError(f"expected a {MESH_DESC_BYTES}-byte mesh descriptor, got {len(desc)} bytes")
        m = cls.__new__(cls)
        m.h = C.c_void_p()
        m.nv = m.nf = 0
        check(lib().gvtb200_mesh_import(C.c_char_p(desc), mode, C.byref(m.h)))
        return m

    def morton(self):
        codes, order = np.zeros(self.nf, np.uint32), np.zeros(self.nf, np.uint32)
        check(lib().gvtb200_mesh_morton(self.h, _fp(codes), _fp(order)))
        return codes, order

    def stats(self):
        st = np.zeros(4, np.uint64)
        check(lib().gvtb200_mesh_stats(self.h, _fp(st)))
        return {"wide_nodes": int(st[0]), "tris": int(st[1]), "depth": int(st[2]), "build_us": int(st[3])}

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        check(lib().gvtb200_intersect(self.h, _fp(rays), len(rays), _fp(hits)))
        return hits

    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.zeros(len(rays), np.int32)
        check(lib().gvtb200_occluded(self.h, _fp(rays), len(rays), _fp(occ)))
        return occ

    def trace(self, rays, m, minv, normi, lights, flags=0, seed=0, begin=0, end=0, cap=None):
        """Adapter::trace.  Returns (rays after the call, moved rays)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE).copy()
        cap = cap if cap is not None else 4 * len(rays) + 16
        moved = np.zeros(cap, RAY_DTYPE)
        n_moved = C.c_size_t(0)
        lights = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        check(
            lib().gvtb200_trace(
                self.h, _fp(rays), len(rays), begin, end, _fp(moved), cap, C.byref(n_moved), _fp(f32(m)), _fp(f32(minv)), _fp(f32(normi)),
                _fp(lights), len(lights), flags, seed,
            )
        )
        return rays, moved[: n_moved.value].copy()


def point_light(pos, color):
    l = np.zeros(1, LIGHT_DTYPE)
    l["type"], l["position"], l["color"] = 0, f32(pos), f32(color)
    return l[0]


def area_light(pos, color, normal, width, height):
    l = np.zeros(1, LIGHT_DTYPE)
    l["type"], l["position"], l["color"], l["normal"], l["width"], l["height"] = 2, f32(pos), f32(color), f32(normal), width, height
    return l[0]


class EngineScene:
    """Thin wrapper of a gvtb200_scene handle (frame-level calls and parity hooks)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    def tlas_order(self, n_inst):
        ids, nn = np.zeros(n_inst, np.int32), C.c_int(0)
        check(lib().gvtb200_scene_tlas_order(self.h, _fp(ids), C.byref(nn)))
        return ids, nn.value

    def camera_rays(self, n):
        rays = np.zeros(n, RAY_DTYPE)
        check(lib().gvtb200_camera_rays(self.h, _fp(rays), n))
        return rays

    def route(self, rays, frm=-1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        nxt, t = np.zeros(len(rays), np.int32), np.zeros(len(rays), np.float32)
        check(lib().gvtb200_route(self.h, _fp(rays), len(rays), frm, _fp(nxt), _fp(t)))
        return nxt, t

    def shuffle(self, rays, frm=-1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out, inst, n = np.zeros(len(rays), RAY_DTYPE), np.zeros(len(rays), np.int32), C.c_size_t(0)
        check(lib().gvtb200_shuffle(self.h, _fp(rays), len(rays), frm, _fp(out), _fp(inst), len(rays), C.byref(n)))
        return out[: n.value].copy(), inst[: n.value].copy()

    def fb_clear(self):
        check(lib().gvtb200_framebuffer_clear(self.h))

    def fb_read_rows(self, host_ptr, tile_rows, index, stride, src_rank=-1):
        """This rank's row tiles of the final image into a full-size host image (asynchronous)."""
        check(lib().gvtb200_framebuffer_read_rows(self.h, C.c_void_p(host_ptr), tile_rows, index, stride, src_rank))

    def set_queue_limit(self, max_rays):
        check(lib().gvtb200_scene_set_queue_limit(self.h, max_rays))

    def fb_clamp(self):
        check(lib().gvtb200_framebuffer_clamp(self.h))

    def fb_read(self, w, h):
        fb = np.zeros((h, w, 4), np.float32)
        check(lib().gvtb200_framebuffer_read(self.h, _fp(fb), fb.size))
        return fb

    def render(self, w, h, row0=0, row1=None, readback=True):
        fb = np.zeros((h, w, 4), np.float32) if readback else None
        ctr = np.zeros(len(COUNTER_FIELDS), np.uint64)
        check(lib().gvtb200_render_frame(self.h, row0, h if row1 is None else row1, _fp(fb), _fp(ctr)))
        return fb, dict(zip(COUNTER_FIELDS, (int(x) for x in ctr)))

    def render_async(self, h, row0=0, row1=None):
        check(lib().gvtb200_render_frame_async(self.h, row0, h if row1 is None else row1))

    def render_tiles(self, h, tile_rows, index, stride, row0=0, row1=None):
        check(lib().gvtb200_render_frame_tiles(self.h, row0, h if row1 is None else row1, tile_rows, index, stride))

    def set_stream(self, cuda_stream):
        check(lib().gvtb200_scene_set_stream(self.h, C.c_void_p(cuda_stream)))

    def render_into(self, host_ptr, h):
        """gvtb200_render_frame into a caller-owned (ideally pinned) host buffer."""
        check(lib().gvtb200_render_frame(self.h, 0, h, C.c_void_p(host_ptr), None))

    def sync(self):
        check(lib().gvtb200_scene_sync(self.h))

    def counters(self):
        ctr = np.zeros(len(COUNTER_FIELDS), np.uint64)
        check(lib().gvtb200_scene_counters(self.h, _fp(ctr)))
        return dict(zip(COUNTER_FIELDS, (int(x) for x in ctr)))

    def profile_counters(self, on=True):
        check(lib().gvtb200_scene_profile_counters(self.h, int(on)))

    def profile_launches(self, on=True):
        check(lib().gvtb200_scene_profile_launches(self.h, 1 if on else 0))

    def launch_records(self):
        """[(kind, ms, counters-after-the-launch)] of the last frame; kind 0 camera+route, 1 closest-hit trace, 2 shade, 3 shadow trace."""
        cap = 64
        n = C.c_int(0)
        kind, ms, ctr = np.zeros(cap, np.int32), np.zeros(cap, np.float32), np.zeros((cap, len(COUNTER_FIELDS)), np.uint64)
        check(lib().gvtb200_scene_launch_records(self.h, cap, C.byref(n), _fp(kind), _fp(ms), _fp(ctr)))
        return [(int(kind[i]), float(ms[i]), dict(zip(COUNTER_FIELDS, (int(x) for x in ctr[i])))) for i in range(n.value)]

    def last_timing(self):
        ms, k = C.c_float(), (C.c_float * 3)()
        check(lib().gvtb200_scene_last_timing(self.h, C.byref(ms), k))
        return ms.value, [k[0], k[1], k[2]]

    def device_framebuffer(self):
        p, st = C.c_void_p(), C.c_void_p()
        check(lib().gvtb200_scene_device_framebuffer(self.h, C.byref(p), C.byref(st)))
        return p.value, st.value

    # ---- Tracer<DomainScheduler> steps ----
    def domain_set_owner(self, owner, rank, world):
        o = np.ascontiguousarray(owner, dtype=np.int32)
        check(lib().gvtb200_domain_set_owner(self.h, _fp(o), len(o), rank, world))
        check(lib().gvtb200_scene_commit(self.h))

    def domain_begin_frame(self):
        check(lib().gvtb200_domain_begin_frame(self.h))

    def domain_drain_local(self):
        check(lib().gvtb200_domain_drain_local(self.h))

    def domain_export(self, world):
        counts, p, n = np.zeros(world, np.uint32), C.c_void_p(), C.c_size_t(0)
        check(lib().gvtb200_domain_export(self.h, _fp(counts), C.byref(p), C.byref(n)))
        return counts, p.value, n.value

    def domain_import(self, d_ptr, n):
        check(lib().gvtb200_domain_import(self.h, C.c_void_p(d_ptr), n))

    def domain_pending(self):
        n = C.c_uint64(0)
        check(lib().gvtb200_domain_pending(self.h, C.byref(n)))
        return n.value

    # ---- NVLink peer path ----
    def p2p_export(self, world):
        """This rank's gvtb200_peer_desc as bytes (to be exchanged with the peers)."""
        buf = C.create_string_buffer(PEER_DESC_BYTES)
        check(lib().gvtb200_p2p_export(self.h, world, buf))
        return buf.raw

    def p2p_attach(self, rank, world, descs):
        """descs: the `world` descriptors concatenated in rank order."""
        descs = bytes(descs)
        if len(descs) != world * PEER_DESC_BYTES:
            raise EngineError(f"expected {world} descriptors of {PEER_DESC_BYTES} bytes, got {len(descs)} bytes")
        check(lib().gvtb200_p2p_attach(self.h, rank, world, C.c_char_p(descs)))

    def p2p_detach(self):
        check(lib().gvtb200_p2p_detach(self.h))

    def image_frame_p2p(self, h, tile_rows=8, root=0, row0=0, row1=None, phases=7):
        check(lib().gvtb200_image_frame_p2p(self.h, row0, h if row1 is None else row1, tile_rows, root, phases))

    def domain_round_p2p(self, rnd):
        """One exchange round; returns the device address of the cumulative exported-ray counter (uint64)."""
        p = C.c_void_p()
        check(lib().gvtb200_domain_round_p2p(self.h, rnd, C.byref(p)))
        return p.value

    def domain_frame_persist(self, root=0, tile_rows=8):
        """One whole DomainScheduler frame as one persistent kernel per rank (asynchronous; sync() reports errors)."""
        check(lib().gvtb200_domain_frame_persist(self.h, root, tile_rows))

    def domain_persist_config(self, grid_ctas=0, shade_sms=0):
        check(lib().gvtb200_domain_persist_config(self.h, grid_ctas, shade_sms))

    def domain_persist_stats(self, n_instances=0):
        """(busy ms of this rank, rays its inbox received, per-instance wide-node visits of a counted frame or None)."""
        busy, rays = C.c_double(0), C.c_uint64(0)
        load = np.zeros(n_instances, np.uint64) if n_instances else None
        check(lib().gvtb200_domain_persist_stats(self.h, C.byref(busy), C.byref(rays), _fp(load), n_instances))
        return busy.value, int(rays.value), load

    def domain_persist_progress(self):
        """[(us since the start barrier, rays retired, rays handed out)] as the root's control warp sampled them."""
        buf, n = np.zeros((256, 3), np.uint32), C.c_int(0)
        check(lib().gvtb200_domain_persist_progress(self.h, _fp(buf), 256, C.byref(n)))
        return buf[: n.value].copy()

    def domain_persist_profile(self):
        """Cycle breakdown of a counted persistent frame (gvtb200_domain_persist_profile), as fractions per role."""
        p = np.zeros(16, np.uint64)
        check(lib().gvtb200_domain_persist_profile(self.h, _fp(p)))
        p = p.astype(np.float64)
        tr, sh = max(1.0, p[1] + p[2] + p[3] + p[4]), max(1.0, p[5] + p[6])
        return {"tracer_warps": int(p[9]), "shader_warps": int(p[8]), "camera_phase_Mcycles_per_warp": p[0] / max(1.0, p[7]) / 1e6,
                "tracer": {"waiting_for_rays": p[1] / tr, "traversal": p[2] / tr, "finishing_rays": p[3] / tr, "tickets": p[4] / tr, "Mcycles_per_warp": tr / max(1.0, p[9]) / 1e6},
                "shader": {"shading": p[5] / sh, "waiting_for_hits": p[6] / sh, "Mcycles_per_warp": sh / max(1.0, p[8]) / 1e6}}

    def domain_set_servers(self, ranks=None, weights=None):
        """ranks: (n_instances, <= 4) serving ranks per instance (-1 = unused), weights: the same shape or None."""
        if ranks is None:
            check(lib().gvtb200_domain_set_servers(self.h, None, None, 0))
            return
        r = np.full((len(ranks), 4), -1, np.int32)
        w = np.zeros((len(ranks), 4), np.float32)
        for i, row in enumerate(ranks):
            for j, x in enumerate(row):
                r[i, j] = x
                w[i, j] = 1.0 if weights is None else weights[i][j]
        check(lib().gvtb200_domain_set_servers(self.h, _fp(r), _fp(w), len(r)))

    def set_instance_mesh(self, instance, mesh):
        check(lib().gvtb200_scene_set_instance_mesh(self.h, instance, mesh.h if mesh is not None else None))
        if mesh is not None:
            self._keep_extra = getattr(self, "_keep_extra", []) + [mesh]

    def commit(self):
        check(lib().gvtb200_scene_commit(self.h))

    def domain_composite_p2p(self, root=0):
        check(lib().gvtb200_domain_composite_p2p(self.h, root))


class GvtScene:
    """The reference's scene API (api::createMesh ... api::render) as a Python object.  The context database behind
    api:: is process-wide, as in the reference, so creating a GvtScene resets it."""

    def __init__(self, flags=0, seed=0, device=None):
        self.A = api()
        if device is not None:
            init(device)
        acheck(self.A.gvtapi_reset())
        acheck(self.A.gvtapi_set_engine_flags(flags, seed))
        self.flags = flags
        self.n_mesh = self.n_inst = self.n_light = 0
        self.W = self.H = 512
        self.samples = 1
        self._have_cam = self._have_film = self._have_ren = False

    def set_flags(self, flags, seed=0):
        self.flags = flags
        acheck(self.A.gvtapi_set_engine_flags(flags, seed))

    def add_mesh(self, verts, tris, one_based=True, kd=None, mtype=0, ks=None, alpha=1.0):
        name = f"mesh{self.n_mesh}".encode()
        self.n_mesh += 1
        v = f32(verts).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)
        acheck(self.A.gvtapi_create_mesh(name))
        acheck(self.A.gvtapi_add_mesh_vertices(name, len(v), _fp(v)))
        if one_based:
            acheck(self.A.gvtapi_add_mesh_triangles(name, len(t), _fp(t)))
        else:
            acheck(self.A.gvtapi_add_mesh_triangles0(name, len(t), _fp(t)))
        if kd is not None:
            acheck(self.A.gvtapi_add_mesh_material(name, mtype, _fp(f32(kd)), _fp(f32(ks if ks is not None else [0.5, 0.5, 0.5])), alpha))
        acheck(self.A.gvtapi_finish_mesh(name, 1))
        return self.n_mesh - 1

    def mesh_face_materials(self, mesh, mats, face_mat):
        """api::addMeshMaterials takes one material per face: expand the (table, index) form."""
        mats = f32(mats).reshape(-1, 8)
        fm = np.asarray(face_mat, dtype=np.int64)
        per = mats[fm]
        types = np.ascontiguousarray(per[:, 0].astype(np.uint32))
        kd, ks, alpha = f32(per[:, 1:4]), f32(per[:, 4:7]), f32(per[:, 7])
        acheck(self.A.gvtapi_add_mesh_materials(f"mesh{mesh}".encode(), len(per), _fp(types), _fp(kd), _fp(ks), _fp(alpha)))

    def mesh_vertex_colors(self, mesh, colors):
        c = f32(colors).reshape(-1, 3)
        acheck(self.A.gvtapi_add_mesh_vertex_color(f"mesh{mesh}".encode(), len(c), _fp(c)))

    def mesh_get(self, mesh, nv):
        name = f"mesh{mesh}".encode()
        nf = C.c_uint(0)
        acheck(self.A.gvtapi_mesh_get(name, C.byref(nf), None, None, None))
        faces, normals, bbox = np.zeros((nf.value, 3), np.uint32), np.zeros((nv, 3), np.float32), np.zeros(6, np.float32)
        acheck(self.A.gvtapi_mesh_get(name, None, _fp(faces), _fp(normals), _fp(bbox)))
        return faces, normals, bbox

    def add_instance(self, mesh, m16):
        name = f"inst{self.n_inst}".encode()
        self.n_inst += 1
        acheck(self.A.gvtapi_add_instance(name, f"mesh{mesh}".encode(), _fp(f32(m16).reshape(16))))
        return self.n_inst - 1

    def instance_get(self, i):
        m, minv, normi, box = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(9, np.float32), np.zeros(6, np.float32)
        acheck(self.A.gvtapi_instance_get(i, _fp(m), _fp(minv), _fp(normi), _fp(box)))
        return m, minv, normi, box

    def add_point_light(self, pos, color):
        acheck(self.A.gvtapi_add_point_light(f"light{self.n_light}".encode(), _fp(f32(pos)), _fp(f32(color))))
        self.n_light += 1

    def add_area_light(self, pos, color, normal, width, height):
        acheck(self.A.gvtapi_add_area_light(f"light{self.n_light}".encode(), _fp(f32(pos)), _fp(f32(color)), _fp(f32(normal)), width, height))
        self.n_light += 1

    def set_camera(self, eye, focus, up, fov, depth=1, samples=1, jitter=0.0):
        self.samples = samples
        fn = self.A.gvtapi_modify_camera if self._have_cam else self.A.gvtapi_add_camera
        acheck(fn(b"Camera", _fp(f32(eye)), _fp(f32(focus)), _fp(f32(up)), fov, depth, samples, jitter))
        self._have_cam = True

    def set_film(self, w, h):
        self.W, self.H = w, h
        fn = self.A.gvtapi_modify_film if self._have_film else self.A.gvtapi_add_film
        acheck(fn(b"Film", w, h, b"gvt_out"))
        self._have_film = True

    def commit(self, schedule=SCHEDULE_IMAGE):
        if not self._have_ren:
            acheck(self.A.gvtapi_add_renderer(b"Renderer", ADAPTER_B200, schedule, b"Camera", b"Film"))
            self._have_ren = True
        return self.engine()

    def engine(self):
        h = self.A.gvtapi_engine_scene(b"Renderer")
        if not h:
            raise EngineError(self.A.gvtapi_last_error().decode())
        return EngineScene(h)

    def set_gpus(self, n):
        """api::setGpus: api::render drives n GPUs of this process (one host thread per GPU)."""
        acheck(self.A.gvtapi_set_gpus(n))

    def set_schedule(self, schedule):
        """scheduler::Image (0) / Domain (1) / AsyncImage (2) / AsyncDomain (3) for the renderer (api::modifyRenderer)."""
        acheck(self.A.gvtapi_modify_renderer(b"Renderer", ADAPTER_B200, schedule, b"Camera", b"Film"))

    def render(self):
        """api::render + the framebuffer (H, W, 4) float32."""
        acheck(self.A.gvtapi_render(b"Renderer"))
        fb = np.zeros((self.H, self.W, 4), np.float32)
        acheck(self.A.gvtapi_framebuffer(_fp(fb), fb.size, None, None))
        return fb

    def writeimage(self, path):
        acheck(self.A.gvtapi_writeimage(b"Renderer", path.encode()))


class StandaloneScene(EngineScene):
    """A gvtb200_scene built directly through the C ABI (no api:: database): several of these can live in one process,
    e.g. one per simulated DomainScheduler rank.  `meshes` are Mesh objects (None = not resident on this rank),
    `instances` is a list of (mesh index, m16, minv16, normi9, box6) as api::addInstance stores them."""

    def __init__(self, meshes, instances, lights, camera, film, flags=0, seed=0, owner=None, rank=0, world=1):
        init(_inited if _inited is not None else 0)
        h = C.c_void_p()
        check(lib().gvtb200_scene_create(C.byref(h)))
        super().__init__(h.value)
        self._keep = meshes
        self.n_inst = len(instances)
        for mi, m, minv, normi, box in instances:
            mesh = meshes[mi]
            box = f32(box)
            check(
                lib().gvtb200_scene_add_instance(
                    self.h, mesh.h if mesh is not None else None, _fp(f32(m)), _fp(f32(minv)), _fp(f32(normi)), _fp(box[:3].copy()), _fp(box[3:].copy())
                )
            )
        ls = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        check(lib().gvtb200_scene_set_lights(self.h, _fp(ls), len(ls)))
        eye, focus, up, fov, depth, samples, jitter = camera
        check(lib().gvtb200_scene_set_camera(self.h, _fp(f32(eye)), _fp(f32(focus)), _fp(f32(up)), fov, depth, samples, jitter))
        self.W, self.H = film
        check(lib().gvtb200_scene_set_film(self.h, self.W, self.H))
        check(lib().gvtb200_scene_set_flags(self.h, flags, seed))
        if owner is not None:
            o = np.ascontiguousarray(owner, dtype=np.int32)
            check(lib().gvtb200_domain_set_owner(self.h, _fp(o), len(o), rank, world))
        check(lib().gvtb200_scene_commit(self.h))

    def close(self):
        if self.h:
            lib().gvtb200_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
