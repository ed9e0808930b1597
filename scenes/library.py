"""Scene descriptions for the BASELINE.json configs and the reference's golden tests.

A scene is a plain description (meshes, instances, lights, camera, film).  `apply(desc, builder)` replays it
through any object that exposes the gvt::render::api-like methods (add_mesh, add_instance, add_point_light,
set_camera, set_film, commit) -- the oracle binding and the product binding both do.
"""
import math
import os
from dataclasses import dataclass, field

import numpy as np

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


@dataclass
class MeshDesc:
    verts: np.ndarray
    tris: np.ndarray
    one_based: bool = True  # api::addMeshTriangles path (1-based + degenerate filter); False = ObjReader path
    kd: tuple = None
    mats: np.ndarray = None  # per-face materials (n,8): type, kd, ks, alpha
    face_mat: np.ndarray = None


@dataclass
class SceneDesc:
    name: str
    meshes: list = field(default_factory=list)
    instances: list = field(default_factory=list)  # (mesh index, m16 column-major)
    lights: list = field(default_factory=list)  # (pos, color)
    eye: tuple = (0, 0, 1)
    focus: tuple = (0, 0, 0)
    up: tuple = (0, 1, 0)
    fov: float = float(np.float32(45.0 * math.pi / 180.0))
    depth: int = 1
    samples: int = 1
    jitter: float = 0.0
    width: int = 512
    height: int = 512
    flags: int = 0

    def n_rays(self):
        return self.width * self.height * self.samples * self.samples


def apply(desc, b):
    ids = []
    for m in desc.meshes:
        mid = b.add_mesh(m.verts, m.tris, one_based=m.one_based, kd=m.kd)
        if m.mats is not None:
            b.mesh_face_materials(mid, m.mats, m.face_mat)
        ids.append(mid)
    for mesh, m16 in desc.instances:
        b.add_instance(ids[mesh], m16)
    for pos, color in desc.lights:
        b.add_point_light(pos, color)
    b.set_camera(desc.eye, desc.focus, desc.up, desc.fov, desc.depth, desc.samples, desc.jitter)
    b.set_film(desc.width, desc.height)
    b.commit()
    return b


def identity():
    return np.eye(4, dtype=np.float32).reshape(16)


def translate_scale(t, s):
    """glm::scale(glm::translate(I, t), s) in fp32, column-major (SimpleApp.cpp:166-169)."""
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = m[1, 1] = m[2, 2] = np.float32(s)
    m[0:3, 3] = np.asarray(t, np.float32)
    return np.ascontiguousarray(m.T).reshape(16)  # column-major


def simple_scene(width=512, height=512):
    """apps/render/SimpleApp.cpp:82-110,163-185,190,209-233 -- golden simple.ppm (25 cone/cube instances)."""
    cone_v = np.array([0.5, 0.0, 0.0, -0.5, 0.5, 0.0, -0.5, 0.25, 0.433013, -0.5, -0.25, 0.43013, -0.5, -0.5, 0.0,
                       -0.5, -0.25, -0.433013, -0.5, 0.25, -0.433013], np.float32).reshape(-1, 3)
    cone_f = np.array([1, 2, 3, 1, 3, 4, 1, 4, 5, 1, 5, 6, 1, 6, 7, 1, 7, 2], np.uint32).reshape(-1, 3)
    cube_v = np.array([-0.5, -0.5, 0.5, 0.5, -0.5, 0.5, 0.5, 0.5, 0.5, -0.5, 0.5, 0.5,
                       -0.5, -0.5, -0.5, 0.5, -0.5, -0.5, 0.5, 0.5, -0.5, -0.5, 0.5, -0.5,
                       0.5, 0.5, 0.5, -0.5, 0.5, 0.5, 0.5, 0.5, -0.5, -0.5, 0.5, -0.5,
                       -0.5, -0.5, 0.5, 0.5, -0.5, 0.5, -0.5, -0.5, -0.5, 0.5, -0.5, -0.5,
                       0.5, -0.5, 0.5, 0.5, 0.5, 0.5, 0.5, -0.5, -0.5, 0.5, 0.5, -0.5,
                       -0.5, -0.5, 0.5, -0.5, 0.5, 0.5, -0.5, -0.5, -0.5, -0.5, 0.5, -0.5], np.float32).reshape(-1, 3)
    cube_f = np.array([1, 2, 3, 1, 3, 4, 17, 19, 20, 17, 20, 18, 6, 5, 8, 6, 8, 7,
                       23, 21, 22, 23, 22, 24, 10, 9, 11, 10, 11, 12, 13, 15, 16, 13, 16, 14], np.uint32).reshape(-1, 3)
    d = SceneDesc("simple", width=width, height=height)
    d.meshes = [MeshDesc(cone_v, cone_f, True, (1.0, 1.0, 1.0)), MeshDesc(cube_v, cube_f, True, (1.0, 1.0, 1.0))]
    inst = 0
    for i in range(-2, 3):
        for j in range(-2, 3):
            m = translate_scale((0.0, i * 0.5, j * 0.5), 0.4)
            d.instances.append((inst % 2, m))  # even -> cone, odd -> cube
            inst += 1
    d.lights = [((1.0, 0.0, -1.0), (1.0, 1.0, 1.0))]
    d.eye, d.focus, d.up = (4.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)
    d.depth, d.samples, d.jitter = 1, 1, 0.5
    return d


def _bunny_camera(d):
    # apps/render/SimpleFileLoadApp.cpp:224-239
    d.lights = [((0.0, 0.1, 0.5), (1.0, 1.0, 1.0))]
    d.eye, d.focus, d.up = (0.0, 0.1, 0.3), (0.0, 0.1, -0.3), (0.0, 1.0, 0.0)
    d.depth, d.samples, d.jitter = 1, 1, 0.0
    return d


def bunny_obj_scene(width=512, height=512):
    """Golden bunny.ppm scene: data/geom/bunny.obj via ObjReader (0-based faces, no filter), identity instance."""
    a = np.load(os.path.join(ASSETS, "bunny_obj.npz"))
    d = SceneDesc("bunny_obj", width=width, height=height)
    d.meshes = [MeshDesc(a["verts"], a["faces"], False, None)]
    d.instances = [(0, identity())]
    return _bunny_camera(d)


def bunny70k_mesh():
    a = np.load(os.path.join(ASSETS, "bunny_70k.npz"))
    return MeshDesc(a["verts"], a["faces"] + 1, True, None)  # through api::addMeshTriangles (1-based)


def bunny_scene(width=1920, height=1080):
    """BASELINE config 1: bun_zipper.ply (69,451 tris), one identity instance, golden-test camera."""
    d = SceneDesc("bunny_70k", width=width, height=height)
    d.meshes = [bunny70k_mesh()]
    d.instances = [(0, identity())]
    return _bunny_camera(d)


def incoherent_scene(width=3840, height=2160):
    """BASELINE config 5: sibenik at 4K, 9 spp (samples = 3: the reference allocates W*H*samples^2 rays, gvtCamera.cpp:205,
    so 8 spp is not expressible), rayMaxDepth 5 = four diffuse bounces: divergent traversal, compaction heavy."""
    d = sibenik_scene(width, height, depth=5, samples=3)
    d.name = "sibenik_incoherent"
    return d


def sibenik_scene(width=1920, height=1080, depth=2, samples=1):
    """BASELINE configs 2/5: sibenik.obj, data/sibenik.conf:2-5 camera; light chosen.
    The eye is inside the cathedral's box, so FLAG_ENTER_INSIDE (deviation F6) is part of the scene, and so is
    FLAG_SHADOW_TO_LIGHT: with the reference's occlusion-to-infinity a closed interior is all shadow."""
    a = np.load(os.path.join(ASSETS, "sibenik.npz"))
    d = SceneDesc("sibenik", width=width, height=height)
    d.meshes = [MeshDesc(a["verts"], a["faces"], False, None, a["mats"], a["face_mat"])]
    d.instances = [(0, identity())]
    d.lights = [((0.0, -8.0, 0.0), (10.0, 10.0, 10.0))]
    d.eye = (-10.7098, -13.9444, 0.299326)
    d.focus = (-0.192988, -12.8887, -0.00787773)
    d.up = (-0.0997908, 0.995004, 0.00291497)
    d.fov = float(np.float32(60.0 * math.pi / 180.0))
    d.depth, d.samples, d.jitter = depth, samples, 0.0
    d.flags = 2 | 8  # FLAG_ENTER_INSIDE | FLAG_SHADOW_TO_LIGHT (closed interior: reference semantics leave it black)
    return d


def bunny_grid_scene(n=32, width=3840, height=2160, pitch=0.2):
    """BASELINE config 3: n x n grid of bunny instances in x-y (two-level instance BVH)."""
    d = SceneDesc(f"bunny_grid_{n}x{n}", width=width, height=height)
    d.meshes = [bunny70k_mesh()]
    for j in range(n):
        for i in range(n):
            d.instances.append((0, translate_scale((i * pitch, j * pitch, 0.0), 1.0)))
    c = (n - 1) * pitch * 0.5
    eye_z = 9.0 * n / 32.0
    d.eye, d.focus, d.up = (c, c + 0.1, eye_z), (c, c + 0.1, 0.0), (0.0, 1.0, 0.0)
    d.lights = [(d.eye, (1.0, 1.0, 1.0))]
    d.depth, d.samples, d.jitter = 1, 1, 0.0
    return d


# ---- config 4: gyroid isosurface, analytic and seed-free ------------------------------------------
_MC_TETS = ((0, 5, 1, 6), (0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6))
_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])


def gyroid_brick(res, brick, bricks=2, extent=8 * math.pi):
    """Marching-tetrahedra triangles of sin x cos y + sin y cos z + sin z cos x = 0 inside brick
    (bx,by,bz) of a bricks^3 split of [0,extent]^3 sampled at `res` cells per axis (whole volume).
    Vertices are not shared between triangles (3 per triangle); returns (verts, tris 1-based)."""
    n = res // bricks
    bx, by, bz = brick
    h = extent / res
    ax = [(np.arange(n + 1) + b * n) * h for b in (bx, by, bz)]
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    F = (np.sin(X) * np.cos(Y) + np.sin(Y) * np.cos(Z) + np.sin(Z) * np.cos(X)).astype(np.float32)
    P = np.stack([X, Y, Z], -1).astype(np.float32)
    out = []
    cs = [(slice(c[0], c[0] + n), slice(c[1], c[1] + n), slice(c[2], c[2] + n)) for c in _CORNERS]
    fv = [F[s].reshape(-1) for s in cs]
    pv = [P[s].reshape(-1, 3) for s in cs]

    def interp(a, b, sel):
        fa, fb = fv[a][sel], fv[b][sel]
        t = (fa / (fa - fb))[:, None]
        return pv[a][sel] + (pv[b][sel] - pv[a][sel]) * t

    for tet in _MC_TETS:
        inside = [fv[k] < 0 for k in tet]
        code = inside[0] * 1 + inside[1] * 2 + inside[2] * 4 + inside[3] * 8
        for c in range(1, 15):
            sel = np.nonzero(code == c)[0]
            if len(sel) == 0:
                continue
            ins = [k for k in range(4) if c >> k & 1]
            outs = [k for k in range(4) if not (c >> k & 1)]
            if len(ins) == 1 or len(ins) == 3:
                a = ins[0] if len(ins) == 1 else outs[0]
                others = outs if len(ins) == 1 else ins
                tri = [interp(tet[a], tet[o], sel) for o in others]
                out.append(np.stack(tri, 1))
            else:
                a, b = ins
                c0, d0 = outs
                p0, p1 = interp(tet[a], tet[c0], sel), interp(tet[a], tet[d0], sel)
                p2, p3 = interp(tet[b], tet[d0], sel), interp(tet[b], tet[c0], sel)
                out.append(np.stack([p0, p1, p2], 1))
                out.append(np.stack([p0, p2, p3], 1))
    tris = np.concatenate(out, 0) if out else np.zeros((0, 3, 3), np.float32)
    verts = tris.reshape(-1, 3).astype(np.float32)
    idx = (np.arange(len(verts), dtype=np.uint32) + 1).reshape(-1, 3)
    return verts, idx


GYROID_RES_100M = 946  # triangles ~ 111.6 * res^2 (measured at res 128 and 256): 946 -> 99.9 M


def gyroid_scene(res=64, bricks=2, width=3840, height=2160, only_bricks=None, brick_fn=None):
    """BASELINE config 4: gyroid isosurface cut into bricks^3 domains, one mesh + identity instance each.
    res=64 is a test size (~0.47 M tris); res = GYROID_RES_100M gives 100 M.  `only_bricks` builds geometry only for the
    listed brick indices (a DomainScheduler rank builds its own bricks); other meshes are tiny placeholders
    with the same bounding box corners so that every rank holds the same instance boxes.  brick_fn: gyroid_brick
    (numpy, default) or gyroid_brick_torch (GPU, for the full-size scene)."""
    brick_fn = brick_fn or gyroid_brick
    d = SceneDesc(f"gyroid_{res}", width=width, height=height)
    extent = 8 * math.pi
    k = 0
    for bz in range(bricks):
        for by in range(bricks):
            for bx in range(bricks):
                if only_bricks is None or k in only_bricks:
                    v, t = brick_fn(res, (bx, by, bz), bricks, extent)
                else:
                    v, t = None, None
                d.meshes.append(MeshDesc(v, t, True, (0.8, 0.8, 0.8)))
                d.instances.append((k, identity()))
                k += 1
    c = extent * 0.5
    r = extent * 1.6
    d.eye = (c + r * 0.577, c + r * 0.577, c + r * 0.577)
    d.focus = (c, c, c)
    d.up = (0.0, 1.0, 0.0)
    d.lights = [(d.eye, (40.0, 40.0, 40.0))]
    d.depth, d.samples, d.jitter = 1, 1, 0.0
    return d


def gyroid_brick_torch(res, brick, bricks=2, extent=8 * math.pi, device="cuda", slab=32):
    """gyroid_brick on the GPU with torch, one z-slab of cells at a time (the 100 M-triangle bricks of BASELINE config 4 are
    out of reach for the numpy version: 3e8 grid points per brick).  Same marching-tetrahedra cases and the same
    interpolation as gyroid_brick; the triangles come out in slab-major order.  Returns numpy (verts, tris 1-based).
    Scene generation only -- this is synthetic input data, not part of the render path."""
    import torch

    n = res // bricks
    bx, by, bz = brick
    h = extent / res
    dd = torch.float64
    ax = [(torch.arange(n + 1, device=device, dtype=dd) + b * n) * h for b in (bx, by)]
    out = []
    for z0 in range(0, n, slab):
        nz = min(slab, n - z0)
        az = (torch.arange(nz + 1, device=device, dtype=dd) + (bz * n + z0)) * h
        X, Y, Z = torch.meshgrid(ax[0], ax[1], az, indexing="ij")
        F = (torch.sin(X) * torch.cos(Y) + torch.sin(Y) * torch.cos(Z) + torch.sin(Z) * torch.cos(X)).to(torch.float32)
        P = torch.stack([X, Y, Z], -1).to(torch.float32)
        del X, Y, Z
        cs = [(slice(c[0], c[0] + n), slice(c[1], c[1] + n), slice(c[2], c[2] + nz)) for c in _CORNERS]
        fv = [F[s].reshape(-1) for s in cs]
        pv = [P[s].reshape(-1, 3) for s in cs]

        def interp(a, b, sel):
            fa, fb = fv[a][sel], fv[b][sel]
            t = (fa / (fa - fb))[:, None]
            return pv[a][sel] + (pv[b][sel] - pv[a][sel]) * t

        for tet in _MC_TETS:
            inside = [(fv[k] < 0).to(torch.int32) for k in tet]
            code = inside[0] + inside[1] * 2 + inside[2] * 4 + inside[3] * 8
            for c in range(1, 15):
                sel = torch.nonzero(code == c).reshape(-1)
                if sel.numel() == 0:
                    continue
                ins = [k for k in range(4) if c >> k & 1]
                outs = [k for k in range(4) if not (c >> k & 1)]
                if len(ins) == 1 or len(ins) == 3:
                    a = ins[0] if len(ins) == 1 else outs[0]
                    others = outs if len(ins) == 1 else ins
                    out.append(torch.stack([interp(tet[a], tet[o], sel) for o in others], 1).cpu())
                else:
                    a, b = ins
                    c0, d0 = outs
                    p0, p1 = interp(tet[a], tet[c0], sel), interp(tet[a], tet[d0], sel)
                    p2, p3 = interp(tet[b], tet[d0], sel), interp(tet[b], tet[c0], sel)
                    out.append(torch.stack([p0, p1, p2], 1).cpu())
                    out.append(torch.stack([p0, p2, p3], 1).cpu())
        del F, P, fv, pv
    tris = torch.cat(out, 0).numpy() if out else np.zeros((0, 3, 3), np.float32)
    verts = np.ascontiguousarray(tris.reshape(-1, 3), dtype=np.float32)
    idx = (np.arange(len(verts), dtype=np.uint32) + 1).reshape(-1, 3)
    return verts, idx
