"""ctypes binding of the CPU oracle (oracle/libgvt_oracle.so) and of the compiled reference pieces
(oracle/_ref/libgvt_ref.so).  TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg, never from gravit_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# actor/Ray.h:68-95 -- 64-byte payload in an 80-byte record
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
assert RAY_DTYPE.itemsize == 80
HIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])

PRIMARY, SHADOW, SECONDARY = 0, 1, 2
FLAG_SMOOTH, FLAG_ENTER_INSIDE, FLAG_BRUTE, FLAG_SHADOW_TO_LIGHT = 1, 2, 4, 8
STATS_FIELDS = (
    "camera_rays routed_initial adapter_calls rays_traced primary_hits shadow_generated shadow_unoccluded "
    "secondary_generated fb_adds rounds rays_sent"
).split()


def build(ref=True):
    """Compile the oracle (and, when /root/reference is present, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", HERE])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libgvt_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.gvto_scene_create.restype = C.c_void_p
        L.gvto_scene_destroy.argtypes = [C.c_void_p]
        L.gvto_add_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_int]
        L.gvto_mesh_material.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float]
        L.gvto_mesh_face_materials.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.gvto_mesh_vertex_colors.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.gvto_mesh_num_faces.argtypes = [C.c_void_p, C.c_int]
        L.gvto_mesh_num_faces.restype = C.c_uint
        L.gvto_mesh_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gvto_add_instance.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.gvto_instance_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.gvto_add_light.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.gvto_clear_lights.argtypes = [C.c_void_p]
        L.gvto_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float]
        L.gvto_set_film.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.gvto_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        L.gvto_set_owner.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.gvto_commit.argtypes = [C.c_void_p]
        L.gvto_tlas_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gvto_camera_rays.argtypes = [C.c_void_p, C.c_void_p]
        L.gvto_camera_rays.restype = C.c_size_t
        L.gvto_route.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.gvto_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.gvto_trace.restype = C.c_size_t
        L.gvto_intersect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
        L.gvto_occluded.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
        L.gvto_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.gvto_route_log.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.gvto_route_log.restype = C.c_size_t
        L.gvto_call_log.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.gvto_call_log.restype = C.c_size_t
        L.gvto_shade.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gvto_morton.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.gvto_fb_to_ppm_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


_ref = None


def ref_lib():
    """The compiled reference pieces, or None when oracle/_ref has not been built (no /root/reference)."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libgvt_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src"):
                build(ref=True)
            else:
                return None
        R = C.CDLL(path)
        R.gvtref_shade.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 5
        R.gvtref_packet_box.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        R.gvtref_mesh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        R.gvtref_instance.argtypes = [C.c_void_p] * 6
        R.gvtref_box_merge_area.argtypes = [C.c_void_p] * 4
        R.gvtref_box_merge_area.restype = C.c_float
        R.gvtref_normal.argtypes = [C.c_void_p] * 3
        _ref = R
    return _ref


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleScene:
    """Mirror of the gvt::render::api scene calls on top of the oracle."""

    def __init__(self, flags=0, seed=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.gvto_scene_create())
        self.flags = flags
        self.L.gvto_set_flags(self.h, flags, seed)
        self.n_inst = 0
        self.W = self.H = 512
        self.samples = 1

    def __del__(self):
        try:
            self.L.gvto_scene_destroy(self.h)
        except Exception:
            pass

    def set_flags(self, flags, seed=0):
        self.flags = flags
        self.L.gvto_set_flags(self.h, flags, seed)

    def add_mesh(self, verts, tris, one_based=True, kd=None, mtype=0, ks=None, alpha=1.0):
        v = f32(verts).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)
        m = self.L.gvto_add_mesh(self.h, _fp(v), len(v), _fp(t), len(t), 1 if one_based else 0)
        if m < 0:
            raise ValueError("vertex index outside bounds")
        if kd is not None:
            kd = f32(kd)
            ksa = f32(ks if ks is not None else [0.5, 0.5, 0.5])
            self.L.gvto_mesh_material(self.h, m, mtype, _fp(kd), _fp(ksa), alpha)
        return m

    def mesh_face_materials(self, mesh, mats, face_mat):
        mats = f32(mats).reshape(-1, 8)
        fm = np.ascontiguousarray(face_mat, dtype=np.int32)
        assert len(fm) == self.mesh_num_faces(mesh)
        self.L.gvto_mesh_face_materials(self.h, mesh, _fp(mats), len(mats), _fp(fm))

    def mesh_vertex_colors(self, mesh, colors):
        c = f32(colors).reshape(-1, 3)
        self.L.gvto_mesh_vertex_colors(self.h, mesh, _fp(c))

    def mesh_num_faces(self, mesh):
        return int(self.L.gvto_mesh_num_faces(self.h, mesh))

    def mesh_get(self, mesh, nv):
        nf = self.mesh_num_faces(mesh)
        faces = np.zeros((nf, 3), np.int32)
        normals = np.zeros((nv, 3), np.float32)
        bbox = np.zeros(6, np.float32)
        self.L.gvto_mesh_get(self.h, mesh, _fp(faces), _fp(normals), _fp(bbox))
        return faces, normals, bbox

    def add_instance(self, mesh, m16):
        m = f32(m16).reshape(16)
        self.n_inst += 1
        return self.L.gvto_add_instance(self.h, mesh, _fp(m))

    def instance_get(self, i):
        m, minv, normi, box = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(9, np.float32), np.zeros(6, np.float32)
        self.L.gvto_instance_get(self.h, i, _fp(m), _fp(minv), _fp(normi), _fp(box))
        return m, minv, normi, box

    def add_point_light(self, pos, color):
        self.L.gvto_add_light(self.h, 0, _fp(f32(pos)), _fp(f32(color)), None, 0.0, 0.0)

    def add_area_light(self, pos, color, normal, width, height):
        self.L.gvto_add_light(self.h, 2, _fp(f32(pos)), _fp(f32(color)), _fp(f32(normal)), width, height)

    def set_camera(self, eye, focus, up, fov, depth=1, samples=1, jitter=0.0):
        self.samples = samples
        self.L.gvto_set_camera(self.h, _fp(f32(eye)), _fp(f32(focus)), _fp(f32(up)), fov, depth, samples, jitter)

    def set_film(self, w, h):
        self.W, self.H = w, h
        self.L.gvto_set_film(self.h, w, h)

    def set_owner(self, owner):
        o = np.ascontiguousarray(owner, dtype=np.int32)
        self.L.gvto_set_owner(self.h, _fp(o), len(o))

    def commit(self):
        self.L.gvto_commit(self.h)

    def tlas_order(self):
        ids = np.zeros(self.n_inst, np.int32)
        nn = C.c_int(0)
        self.L.gvto_tlas_get(self.h, _fp(ids), C.byref(nn))
        return ids, nn.value

    def camera_rays(self):
        rays = np.zeros(self.W * self.H * self.samples * self.samples, RAY_DTYPE)
        n = self.L.gvto_camera_rays(self.h, _fp(rays))
        assert n == len(rays)
        return rays

    def route(self, rays, frm=-1, simd=8):
        rays = np.ascontiguousarray(rays)
        nxt = np.zeros(len(rays), np.int32)
        t = np.zeros(len(rays), np.float32)
        self.L.gvto_route(self.h, _fp(rays), len(rays), frm, simd, _fp(nxt), _fp(t))
        return nxt, t

    def trace(self, inst, rays, threads=1, cap=None):
        """Adapter::trace: returns (rays mutated in place copy, moved, first-hit records)."""
        rays = np.ascontiguousarray(rays).copy()
        cap = cap if cap is not None else len(rays) * 4 + 16
        moved = np.zeros(cap, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        n = self.L.gvto_trace(self.h, inst, _fp(rays), len(rays), _fp(moved), cap, _fp(hits), threads)
        if n == C.c_size_t(-1).value:
            raise RuntimeError("moved capacity too small")
        return rays, moved[:n].copy(), hits

    def intersect(self, mesh, rays, brute=False, threads=1):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), HIT_DTYPE)
        self.L.gvto_intersect(self.h, mesh, _fp(rays), len(rays), int(brute), _fp(hits), threads)
        return hits

    def occluded(self, mesh, rays, brute=False, threads=1):
        rays = np.ascontiguousarray(rays)
        occ = np.zeros(len(rays), np.int32)
        self.L.gvto_occluded(self.h, mesh, _fp(rays), len(rays), int(brute), _fp(occ), threads)
        return occ

    def render(self, schedule=0, world_size=1, threads=1):
        fb = np.zeros((self.H, self.W, 4), np.float32)
        st = np.zeros(len(STATS_FIELDS), np.int64)
        self.L.gvto_render(self.h, schedule, world_size, threads, _fp(fb), _fp(st))
        return fb, dict(zip(STATS_FIELDS, (int(x) for x in st)))

    def route_log(self, world_size):
        n = self.L.gvto_route_log(self.h, None, 0)
        out = np.zeros(n, np.int64)
        self.L.gvto_route_log(self.h, _fp(out), n)
        return out.reshape(-1, world_size, world_size)

    def call_log(self):
        n = self.L.gvto_call_log(self.h, None, 0)
        out = np.zeros(n, np.int32)
        self.L.gvto_call_log(self.h, _fp(out), n)
        return out.reshape(-1, 3)

    def morton(self, mesh):
        codes = np.zeros(self.mesh_num_faces(mesh), np.uint32)
        self.L.gvto_morton(self.h, mesh, _fp(codes))
        return codes


def fb_to_ppm_bytes(fb):
    """IceTComposite::write byte stream (without the 15-byte header)."""
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    h, w = fb.shape[:2]
    out = np.zeros(h * w * 3, np.uint8)
    lib().gvto_fb_to_ppm_bytes(_fp(fb), w, h, _fp(out))
    return out


def read_ppm(data):
    """Parse a P6 file (bytes) -> (w, h, uint8 array of h*w*3 in file order)."""
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6"
    w, h = (int(x) for x in parts[1].split())
    assert parts[2] == b"255"
    return w, h, np.frombuffer(parts[3], np.uint8, w * h * 3)


def image_diff(a, b):
    """apps/render/ImageDiff.cpp:44-66: sum of absolute byte differences."""
    return int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum())


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 999.0 if mse == 0 else 10 * np.log10(255.0**2 / mse)
