"""Parity of the CUDA path against the CPU oracle, through the C ABI (include/gvtb200.h).  Needs a B200.

Bars (BASELINE.json north_star): Morton codes, routing and queue contents bit-exact; hit ids >= 99.9 % (here: all equal
on these inputs), t within 1e-5 relative (here: bit-exact, the triangle test is the same IEEE sequence on both sides);
images >= 50 dB PSNR against the oracle framebuffer and within the reference's own tolerance of its golden PPMs."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from scenes import library
from tests.scene_util import golden_ppm, grazing_rays, oracle_scene, random_rays, sorted_payload

pytestmark = pytest.mark.gpu


def _pair(g, desc, flags=None):
    flags = desc.flags if flags is None else flags
    o = oracle_scene(desc, flags)
    s = library.apply(desc, g.GvtScene(flags=flags))
    return o, s, s.engine()


def _mesh_handle(g, o, desc, k=0, **kw):
    nv = len(np.asarray(desc.meshes[k].verts).reshape(-1, 3))
    faces, normals, bbox = o.mesh_get(k, nv)
    return g.Mesh(desc.meshes[k].verts, faces, normals=normals, **kw), bbox


SCENES = {
    "simple": lambda: library.simple_scene(128, 128),
    "bunny_obj": lambda: library.bunny_obj_scene(256, 256),
    "bunny70k": lambda: library.bunny_scene(320, 180),
    "sibenik": lambda: library.sibenik_scene(192, 108),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_blas_morton_and_hits_bit_exact(gpu, name):
    g = gpu
    d = SCENES[name]()
    o = oracle_scene(d)
    for k in range(len(d.meshes)):
        mesh, bbox = _mesh_handle(g, o, d, k)
        codes, order = mesh.morton()
        assert np.array_equal(codes, o.morton(k))
        assert np.array_equal(np.sort(order), np.arange(mesh.nf))
        assert np.all(np.diff(codes[order].astype(np.int64)) >= 0)
        # stable: equal keys keep input order
        same = codes[order][1:] == codes[order][:-1]
        assert np.all(order[1:][same] > order[:-1][same])
        assert mesh.stats()["tris"] == mesh.nf
        rays = np.concatenate([o.camera_rays(), random_rays(bbox, 20000, seed=k + 7)])
        ho = o.intersect(k, rays, threads=8)
        hg = mesh.intersect(rays)
        assert np.array_equal(ho, hg)
        assert np.array_equal(o.occluded(k, rays, threads=8), mesh.occluded(rays))
        mesh.close()


@pytest.mark.parametrize("name", ["simple", "bunny_obj"])
def test_blas_grazing_rays_bit_exact(gpu, name):
    """Hits just above RAY_EPSILON (ray origins 1e-6 .. 1e-4 off a face) survive the quantised child boxes; also through
    the trace kernel's lane-refill / triangle-postponing path (Adapter::trace)."""
    g = gpu
    d = SCENES[name]()
    o = oracle_scene(d)
    for k in range(len(d.meshes)):
        mesh, bbox = _mesh_handle(g, o, d, k)
        v = po.f32(d.meshes[k].verts).reshape(-1, 3)
        faces, _, _ = o.mesh_get(k, len(v))
        rays = np.concatenate([grazing_rays(v, faces, reps=max(20, 40000 // max(1, len(faces)))), random_rays(bbox, 200000, seed=104 + 4 * k)])
        ho = o.intersect(k, rays, threads=8)
        assert np.array_equal(ho, mesh.intersect(rays))
        assert np.array_equal(o.occluded(k, rays, threads=8), mesh.occluded(rays))
        # the same rays through k_trace (adapter mode, no lights): rays that miss come back, in any order
        ident = np.eye(4, dtype=np.float32).reshape(16)
        normi = np.eye(3, dtype=np.float32).reshape(9)
        _, moved = mesh.trace(rays, ident, ident, normi, np.zeros(0, g.LIGHT_DTYPE), flags=0)
        assert sorted(moved["id"].tolist()) == sorted(rays["id"][ho["prim"] < 0].tolist())
        mesh.close()


def test_blas_edge_cases(gpu):
    g = gpu
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    rays = np.zeros(4, g.RAY_DTYPE)
    rays["origin"] = [[0.2, 0.2, 1], [0.2, 0.2, -1], [2, 2, 1], [0.25, 0.25, 1]]
    rays["direction"] = [[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1]]  # axis-parallel: 1/0 in the slab test
    for faces, expect in (
        (np.zeros((0, 3), np.uint32), [-1, -1, -1, -1]),
        (np.array([[0, 1, 2]], np.uint32), [0, 0, -1, 0]),
        (np.array([[0, 1, 2], [0, 1, 2], [0, 1, 2]], np.uint32), [0, 0, -1, 0]),  # t tie -> smallest primitive id
        (np.array([[0, 1, 2], [1, 3, 2]], np.uint32), [0, 0, -1, 0]),  # shared edge
    ):
        m = g.Mesh(v, faces)
        assert m.intersect(rays)["prim"].tolist() == expect
        assert m.intersect(rays[:0]).shape == (0,)
        m.close()
    with pytest.raises(g.EngineError):
        g.Mesh(v, np.array([[0, 1, 9]], np.uint32))


def _hostile_mesh(kind):
    rng = np.random.default_rng(42)
    if kind == "soup":
        c = rng.random((3000, 1, 3)).astype(np.float32)
        tri = c + (rng.random((3000, 3, 3)).astype(np.float32) - 0.5) * 0.2
    elif kind == "slivers_and_dust":
        a = rng.random((300, 3)).astype(np.float32)
        dd = rng.normal(size=(300, 3)).astype(np.float32)
        sl = np.stack([a, a + dd * 0.8, a + dd * 0.8 + rng.normal(size=(300, 3)).astype(np.float32) * 0.01], axis=1)
        c = rng.random((3000, 1, 3)).astype(np.float32)
        dust = c + (rng.random((3000, 3, 3)).astype(np.float32) - 0.5) * 0.01
        tri = np.concatenate([sl, dust]).astype(np.float32)
    else:
        one = np.array([[0.1, 0.1, 0.5], [0.9, 0.2, 0.5], [0.4, 0.8, 0.5]], np.float32)
        tri = np.repeat(one[None], 700, axis=0)
    verts = tri.reshape(-1, 3)
    d = library.SceneDesc(kind, width=32, height=32)
    d.meshes.append(library.MeshDesc(verts, np.arange(len(verts), dtype=np.uint32).reshape(-1, 3) + 1, True, None))
    d.instances.append((0, np.eye(4, dtype=np.float32).T.reshape(-1)))
    d.lights.append(((0.5, 0.5, 2.0), (1.0, 1.0, 1.0)))
    d.eye, d.focus = (0.5, 0.5, 3.0), (0.5, 0.5, 0.0)
    return d


@pytest.mark.parametrize("kind", ["soup", "slivers_and_dust", "coincident"])
def test_blas_build_on_hostile_meshes(gpu, kind):
    """The device build (binned SAH, insertion-based optimisation kernels k_ri_*, optimal widening) on a triangle soup, long
    slivers mixed with dust, and 700 copies of one triangle: every triangle referenced once, hits equal the oracle's bit for
    bit, twice in a row (the build scratch comes from a pool)."""
    g = gpu
    d = _hostile_mesh(kind)
    o = oracle_scene(d)
    for rep in range(2):
        mesh, bbox = _mesh_handle(g, o, d, 0)
        assert mesh.stats()["tris"] == mesh.nf
        rays = np.concatenate([o.camera_rays(), random_rays(bbox, 20000, seed=9 + rep)])
        ho = o.intersect(0, rays, threads=8)
        assert np.array_equal(ho, mesh.intersect(rays))
        assert np.array_equal(o.occluded(0, rays, threads=8), mesh.occluded(rays))
        mesh.close()


@pytest.mark.parametrize("name,samples", [("simple", 1), ("bunny_obj", 2), ("sibenik", 1)])
def test_camera_rays_bit_exact(gpu, name, samples):
    d = SCENES[name]()
    d.samples = samples
    d.jitter = 1.5 if samples > 1 else d.jitter  # int truncation -> window 1 (gvtCamera.h:136)
    o, s, e = _pair(gpu, d)
    a, b = o.camera_rays(), e.camera_rays(d.n_rays())
    assert np.array_equal(a.view(np.uint8).reshape(-1, 80)[:, :64], b.view(np.uint8).reshape(-1, 80)[:, :64])


@pytest.mark.parametrize("name", ["simple", "grid"])
def test_tlas_and_routing_bit_exact(gpu, name):
    d = library.simple_scene(160, 160) if name == "simple" else library.bunny_grid_scene(6, 200, 120)
    o, s, e = _pair(gpu, d)
    ids_o, nn_o = o.tlas_order()
    ids_g, nn_g = e.tlas_order(len(d.instances))
    assert np.array_equal(ids_o, ids_g) and nn_o == nn_g
    rays = o.camera_rays()
    no, to = o.route(rays, -1)
    ng, tg = e.route(rays, -1)
    assert np.array_equal(no, ng) and np.array_equal(to.view(np.uint32), tg.view(np.uint32))
    # second hop: rays that left instance k (`from` exclusion, BVH.h:112)
    k = int(np.bincount(no[no >= 0]).argmax())
    sub = rays[no == k].copy()
    sub["origin"] += sub["direction"] * (to[no == k] * np.float32(0.95))[:, None]
    no2, to2 = o.route(sub, k)
    ng2, tg2 = e.route(sub, k)
    assert np.array_equal(no2, ng2) and np.array_equal(to2.view(np.uint32), tg2.view(np.uint32))
    assert np.all(no2 != k)
    # shuffleRays: advance + group by instance; contents compared as multisets
    out, inst = e.shuffle(rays, -1)
    assert np.array_equal(np.bincount(inst, minlength=len(d.instances)), np.bincount(no[no >= 0], minlength=len(d.instances)))
    adv = rays[no >= 0].copy()
    adv["origin"] = adv["origin"] + adv["direction"] * (to[no >= 0] * np.float32(0.95))[:, None]
    assert np.array_equal(sorted_payload(out), sorted_payload(adv))


@pytest.mark.parametrize("name,flags", [("simple", 0), ("simple", 1), ("bunny_obj", 1), ("bunny70k", 0)])
def test_adapter_trace_queue_contents_bit_exact(gpu, name, flags):
    """Adapter::trace (Adapter.h:82-84): moved_rays and the in-place state of rayList, as multisets of 64-byte payloads."""
    g = gpu
    d = SCENES[name]()
    o = oracle_scene(d, flags)
    rays = o.camera_rays()
    nxt, t = o.route(rays, -1)
    inst = int(np.bincount(nxt[nxt >= 0]).argmax())
    q = rays[nxt == inst].copy()
    q["origin"] = q["origin"] + q["direction"] * (t[nxt == inst] * np.float32(0.95))[:, None]
    mesh_idx = d.instances[inst][0]
    mesh, _ = _mesh_handle(g, o, d, mesh_idx, material=(0, tuple(d.meshes[mesh_idx].kd or (0.5, 0.5, 0.5)), (0.5, 0.5, 0.5), 1.0))
    m, minv, normi, box = o.instance_get(inst)
    lights = np.array([g.point_light(p, c) for p, c in d.lights], g.LIGHT_DTYPE)
    ro, mo, ho = o.trace(inst, q)
    rg, mg = mesh.trace(q, m, minv, normi, lights, flags=flags)
    assert len(mo) == len(mg)
    assert np.array_equal(sorted_payload(mo), sorted_payload(mg))
    assert np.array_equal(ro.view(np.uint8).reshape(-1, 80)[:, :64], rg.view(np.uint8).reshape(-1, 80)[:, :64])
    # a sub-range call touches only [begin, end) (HeterogeneousMeshAdapter.cpp:69-79)
    half = len(q) // 2
    rg2, mg2 = mesh.trace(q, m, minv, normi, lights, flags=flags, begin=half, end=len(q))
    assert np.array_equal(rg2[:half], q[:half])
    assert np.array_equal(rg2.view(np.uint8).reshape(-1, 80)[half:, :64], ro.view(np.uint8).reshape(-1, 80)[half:, :64])
    # shadow rays arriving from another instance: occluded ones vanish, the rest are handed back unchanged
    sh = mo[mo["type"] == 1][:5000]
    if len(sh):
        ro3, mo3, _ = o.trace(inst, sh)
        rg3, mg3 = mesh.trace(sh, m, minv, normi, lights, flags=flags)
        assert np.array_equal(sorted_payload(mo3), sorted_payload(mg3))
    # capacity error instead of a crash
    with pytest.raises(g.EngineError, match="capacity"):
        mesh.trace(q, m, minv, normi, lights, flags=flags, cap=1)
    mesh.close()


@pytest.mark.parametrize("name", ["simple", "bunny70k"])
def test_reference_side_adapter_runs_for_real(gpu, name):
    """INTEGRATION.md section 1 executed: gravit_b200/host/gvt_shim/B200MeshAdapter.h compiled against GraviT's OWN headers and
    linked with GraviT's own Ray.cpp / Mesh.cpp / Light.cpp / Material.cpp / BBox.cpp (oracle/_ref, tests/ref_adapter/harness.cpp):
    a real gvt::render::data::primitives::Mesh, a real RayVector and a real PointLight go through Adapter::trace
    (Adapter.h:82-84, EmbreeMeshAdapter.cpp:625-660) on the GPU.  moved_rays and the in-place state of rayList equal the
    oracle's byte for byte (64-byte payloads; moved_rays as a multiset: the reference's order is nondeterministic)."""
    import ctypes as C

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libgvt_ref_adapter.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libgvt_ref_adapter.so not built (needs /root/reference at build time)")
    H = C.CDLL(path)
    H.refadapter_last_error.restype = C.c_char_p
    H.refadapter_trace.restype = C.c_long
    H.refadapter_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    d = SCENES[name]()
    o = oracle_scene(d, 0)  # HEAD's flat shading: what the binding asks for
    rays = o.camera_rays()
    nxt, t = o.route(rays, -1)
    inst = int(np.bincount(nxt[nxt >= 0]).argmax())
    q = rays[nxt == inst].copy()
    q["origin"] = q["origin"] + q["direction"] * (t[nxt == inst] * np.float32(0.95))[:, None]
    mi, m16 = d.instances[inst]
    mesh = d.meshes[mi]
    assert mesh.one_based  # the harness feeds Mesh::addFace (1-based, degenerate faces dropped)
    ro, mo, _ = o.trace(inst, q)
    verts = np.ascontiguousarray(mesh.verts, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(mesh.tris, np.uint32).reshape(-1, 3)
    kd = np.asarray(mesh.kd or (0.5, 0.5, 0.5), np.float32)
    lp = np.ascontiguousarray([p for p, _ in d.lights], np.float32)
    lc = np.ascontiguousarray([c for _, c in d.lights], np.float32)
    for begin, end in ((0, 0), (len(q) // 3, len(q))):  # end == 0 means "all" (EmbreeMeshAdapter.cpp:642); then a sub-range
        rg = q.copy()
        cap = 4 * len(q) + 16
        mg = np.zeros(cap, gpu.RAY_DTYPE)
        n = H.refadapter_trace(verts.ctypes.data, len(verts), tris.ctypes.data, len(tris), kd.ctypes.data, rg.ctypes.data, len(rg), begin, end,
                               np.ascontiguousarray(m16, np.float32).ctypes.data, lp.ctypes.data, lc.ctypes.data, len(lp), mg.ctypes.data, cap)
        assert n >= 0, H.refadapter_last_error()
        mg = mg[:n]
        if begin == 0:
            assert n == len(mo) and np.array_equal(sorted_payload(mo), sorted_payload(mg))
            assert np.array_equal(ro.view(np.uint8).reshape(-1, 80)[:, :64], rg.view(np.uint8).reshape(-1, 80)[:, :64])
        else:
            ro2, mo2, _ = o.trace(inst, q[begin:])
            assert n == len(mo2) and np.array_equal(sorted_payload(mo2), sorted_payload(mg))
            assert np.array_equal(rg[:begin], q[:begin])  # rays outside [begin, end) are not touched (HeterogeneousMeshAdapter.cpp:69-79)
            assert np.array_equal(ro2.view(np.uint8).reshape(-1, 80)[:, :64], rg.view(np.uint8).reshape(-1, 80)[begin:, :64])


@pytest.mark.parametrize("name,flags", [("simple", 1), ("simple", 0), ("bunny_obj", 1), ("bunny70k", 0)])
def test_frame_matches_oracle_and_golden(gpu, name, flags):
    full = {"simple": lambda: library.simple_scene(), "bunny_obj": lambda: library.bunny_obj_scene(), "bunny70k": lambda: library.bunny_scene(512, 512)}
    d = full[name]()
    o, s, e = _pair(gpu, d, flags)
    ref, st = o.render(threads=8)
    fb = s.render()
    ctr = e.counters()
    for k in ("camera_rays", "routed_initial", "primary_hits", "shadow_generated", "shadow_unoccluded", "fb_adds"):
        assert ctr[k] == st[k], (k, ctr[k], st[k])
    assert ctr["instance_visits"] == st["rays_traced"] + st["shadow_generated"]
    assert np.array_equal(fb, ref)  # one light, one sample: a single add per pixel, bit-exact
    a, b = po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)
    assert po.image_diff(a, b) == 0
    if flags == 1 and name in ("simple", "bunny_obj"):  # the reference's own fixtures (smooth normals)
        _, _, gold = golden_ppm("simple" if name == "simple" else "bunny")
        assert po.image_diff(a, gold) < 300  # CMakeLists.txt:666
        assert po.psnr(a, gold) >= 50.0


def test_frame_multi_light_multi_sample(gpu):
    d = library.simple_scene(200, 200)
    d.lights = d.lights + [((2.0, 2.0, 1.0), (0.6, 0.5, 0.4)), ((3.0, -1.0, 0.5), (0.2, 0.3, 0.9))]
    d.samples, d.jitter = 2, 1.0
    o, s, e = _pair(gpu, d, 1)
    ref, st = o.render(threads=8)
    fb = s.render()
    ctr = e.counters()
    for k in ("primary_hits", "shadow_generated", "shadow_unoccluded", "fb_adds"):
        assert ctr[k] == st[k]
    assert np.allclose(fb, ref, atol=2e-6)  # several adds per pixel: fp32 reassociation only
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) >= 50.0


@pytest.mark.parametrize("flags", [2, 2 | 8])
def test_frame_sibenik_inside_box_and_bounces(gpu, flags):
    """Config 2: camera inside the instance box (deviation F6), per-face materials, shadow + 1 bounce; with the
    reference's occlusion-to-infinity (flags 2: the closed interior is all shadow) and with FLAG_SHADOW_TO_LIGHT."""
    d = library.sibenik_scene(384, 216, depth=2)
    d.flags = flags
    o, s, e = _pair(gpu, d)
    ref, st = o.render(threads=8)
    fb = s.render()
    ctr = e.counters()
    assert ctr["routed_initial"] == st["routed_initial"] == d.n_rays()
    assert abs(ctr["primary_hits"] - st["primary_hits"]) <= 1e-3 * st["primary_hits"]
    assert abs(ctr["secondary_generated"] - st["secondary_generated"]) <= 1e-3 * st["secondary_generated"] + 2  # same counter-based RNG
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) >= 50.0
    assert abs(ctr["shadow_unoccluded"] - st["shadow_unoccluded"]) <= 1e-3 * st["shadow_unoccluded"] + 2
    if flags & 8:
        assert st["shadow_unoccluded"] > 0.2 * st["shadow_generated"]
    d1 = library.sibenik_scene(384, 216, depth=1)
    d1.flags = flags
    o1, s1, e1 = _pair(gpu, d1)
    ref1, _ = o1.render(threads=8)
    assert np.array_equal(s1.render(), ref1)


def test_materials_and_area_light(gpu):
    """Phong / Blinn / per-vertex colour / area light."""
    g = gpu
    d = library.bunny_obj_scene(200, 200)
    for mtype in (1, 2):
        o = po.OracleScene(flags=1)
        s = g.GvtScene(flags=1)
        for b in (o, s):
            m = b.add_mesh(d.meshes[0].verts, d.meshes[0].tris, one_based=False, kd=(0.6, 0.5, 0.4), mtype=mtype, ks=(0.5, 0.4, 0.3), alpha=12.0)
            b.add_instance(m, library.identity())
            b.add_point_light(*d.lights[0])
            b.add_area_light((0.1, 0.3, 0.4), (0.5, 0.5, 0.5), (0.0, -1.0, 0.0), 0.1, 0.1)
            b.set_camera(d.eye, d.focus, d.up, d.fov, 1, 1, 0.0)
            b.set_film(d.width, d.height)
            b.commit()
        ref, st = o.render(threads=8)
        fb = s.render()
        assert s.engine().counters()["shadow_generated"] == st["shadow_generated"]
        assert np.allclose(fb, ref, atol=1e-4)  # powf differs by a few ulp between libm and the device
        assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) >= 50.0
    # per-vertex colours
    o = po.OracleScene(flags=1)
    s = g.GvtScene(flags=1)
    nv = len(d.meshes[0].verts)
    col = np.random.default_rng(3).random((nv, 3)).astype(np.float32)
    for b in (o, s):
        m = b.add_mesh(d.meshes[0].verts, d.meshes[0].tris, one_based=False)
        b.mesh_vertex_colors(m, col)
        b.add_instance(m, library.identity())
        b.add_point_light(*d.lights[0])
        b.set_camera(d.eye, d.focus, d.up, d.fov, 1, 1, 0.0)
        b.set_film(d.width, d.height)
        b.commit()
    assert np.array_equal(s.render(), o.render(threads=8)[0])


def test_full_size_properties(gpu):
    """BASELINE config 1 at 1920x1080: size-independent properties instead of an oracle run."""
    d = library.bunny_scene(1920, 1080)
    s = library.apply(d, gpu.GvtScene(flags=0))
    e = s.engine()
    fb, c = e.render(1920, 1080)
    assert c["camera_rays"] == 1920 * 1080
    assert 0 < c["shadow_generated"] <= c["primary_hits"]  # one light; Shade() skips hits whose normal faces away
    assert c["fb_adds"] == c["shadow_unoccluded"] <= c["shadow_generated"]
    assert c["instance_visits"] == c["routed_initial"] + c["shadow_generated"]
    assert int((fb[..., 3] > 0).sum()) == c["fb_adds"]  # one contribution per lit pixel
    assert fb[..., :3].max() <= 1.0 and fb.min() >= 0.0
    fb2, c2 = e.render(1920, 1080)
    assert np.array_equal(fb, fb2) and c == c2  # idempotent / deterministic
    # row sharding (ImageScheduler): two half frames sum to the full frame
    top, _ = e.render(1920, 1080, 0, 540)
    bot, _ = e.render(1920, 1080, 540, 1080)
    assert np.array_equal(top + bot, fb) and not top[540:].any() and not bot[:540].any()
    # the 512x512 oracle-checked frame is a crop-independent check of the same mesh: hit fraction is similar
    assert 0.05 < c["primary_hits"] / c["camera_rays"] < 0.5


def test_incoherent_config4_reduced_film(gpu):
    """BASELINE configs[4] at a reduced film: sibenik, samples = 3 (W*H*samples^2 camera rays, gvtCamera.cpp:205), rayMaxDepth 5
    (the bounce loop of EmbreeMeshAdapter.cpp:584-598 run four times).  The random numbers of a bounce are a function of (ray id,
    depth) on both sides (the reference's per-TBB-chunk RandEngine, core/math/RandEngine.h:43-81, is not reproducible and is
    replaced on both sides: a documented deviation), so the counters agree to 1e-3 and the image to >= 50 dB."""
    d = library.incoherent_scene(240, 135)
    assert d.depth == 5 and d.samples == 3
    o, s, e = _pair(gpu, d)
    ref, st = o.render(threads=8)
    fb = s.render()
    ctr = e.counters()
    assert ctr["camera_rays"] == st["camera_rays"] == 240 * 135 * 9
    assert st["secondary_generated"] > 0.05 * st["camera_rays"]  # the bounces are really there
    for k in ("primary_hits", "shadow_generated", "shadow_unoccluded", "secondary_generated", "fb_adds"):
        assert abs(ctr[k] - st[k]) <= 1e-3 * st[k] + 2, (k, ctr[k], st[k])
    assert ctr["queue_overflows"] == 0 and ctr["stack_overflows"] == 0
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) >= 50.0
    assert np.allclose(fb, ref, atol=2e-3)


def test_instanced_grid_frame_bit_identical(gpu):
    """BASELINE configs[2] in small: a 9x9 grid of bunny.obj instances (two-level BVH: 81 instance boxes over one BLAS, rays
    that miss one instance are routed on to the next, TracerBase.h:392-401).  Depth 1, one light: one add per pixel, so the
    framebuffer is bit-identical to the oracle's, and so are the routing / hit / shadow counters."""
    a = np.load(os.path.join(library.ASSETS, "bunny_obj.npz"))
    d = library.SceneDesc("grid9", width=480, height=360)
    d.meshes = [library.MeshDesc(a["verts"], a["faces"], False, None)]
    n, pitch = 9, 0.2
    for j in range(n):
        for i in range(n):
            d.instances.append((0, library.translate_scale((i * pitch, j * pitch, 0.0), 1.0)))
    c = (n - 1) * pitch * 0.5
    d.eye, d.focus, d.up = (c + 0.3, c + 0.1, 2.2), (c, c + 0.1, 0.0), (0.0, 1.0, 0.0)  # slightly oblique: rays cross several boxes
    d.lights = [(d.eye, (1.0, 1.0, 1.0))]
    o, s, e = _pair(gpu, d, 1)
    ref, st = o.render(threads=8)
    fb = s.render()
    ctr = e.counters()
    for k in ("camera_rays", "routed_initial", "primary_hits", "shadow_generated", "shadow_unoccluded", "fb_adds"):
        assert ctr[k] == st[k], (k, ctr[k], st[k])
    assert ctr["instance_visits"] == st["rays_traced"] + st["shadow_generated"] > ctr["routed_initial"] + ctr["shadow_generated"]  # rays did move on
    assert np.array_equal(fb, ref)
    # the persistent scheduler renders the same frame
    e.domain_frame_persist(0, 8)
    e.sync()
    assert np.array_equal(e.fb_read(d.width, d.height), ref)


@pytest.mark.parametrize("name", ["bunny", "sibenik"])
def test_full_1080p_frames_against_the_oracle(gpu, name):
    """The 1080p frames of the north-star target (bunny, primary + shadow: bit-identical) and of BASELINE configs[1] (sibenik,
    shadow + 1 bounce: >= 50 dB, counters to 1e-3) against the CPU oracle at FULL size, once."""
    d = library.bunny_scene(1920, 1080) if name == "bunny" else library.sibenik_scene(1920, 1080, depth=2)
    o, s, e = _pair(gpu, d, 0 if name == "bunny" else None)
    ref, st = o.render(threads=os.cpu_count() or 8)
    fb = s.render()
    ctr = e.counters()
    assert ctr["camera_rays"] == st["camera_rays"] == 1920 * 1080
    if name == "bunny":
        for k in ("routed_initial", "primary_hits", "shadow_generated", "shadow_unoccluded", "fb_adds"):
            assert ctr[k] == st[k], (k, ctr[k], st[k])
        assert np.array_equal(fb, ref)
    else:
        for k in ("primary_hits", "shadow_generated", "shadow_unoccluded", "secondary_generated", "fb_adds"):
            assert abs(ctr[k] - st[k]) <= 1e-3 * st[k] + 2, (k, ctr[k], st[k])
        assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) >= 50.0


def _standalone_ranks(g, d, flags, world):
    """One engine scene per simulated rank on the same GPU: rank r holds only the meshes of the instances it owns."""
    from gravit_b200 import schedulers

    ranks = []
    for r in range(world):
        scene, owner = schedulers.build_domain_rank(d, r, world, flags=flags)
        ranks.append(scene)
    return ranks, owner


@pytest.mark.parametrize("world", [2, 3])
def test_domain_scheduler_simulated_ranks(gpu, world):
    """Tracer<DomainScheduler> (DomainTracer.h:185-496) with P ranks simulated on one GPU: per-round ray-routing count
    matrices equal the oracle's bit for bit, and the composited image equals the single-process image."""
    import torch

    g = gpu
    d = library.simple_scene(256, 256)
    flags = 1
    o = oracle_scene(d, flags)
    o.set_owner(np.arange(len(d.instances)) % world)
    ref, st = o.render(schedule=1, world_size=world, threads=8)
    log_o = o.route_log(world)
    ref1, _ = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(g, d, flags, world)
    for e in ranks:
        e.domain_begin_frame()
    log_g = []
    for rnd in range(100):
        sends = []
        for e in ranks:
            e.domain_drain_local()
            counts, ptr, n = e.domain_export(world)
            # copy the wire records out before the next export reuses the buffer
            buf = torch.empty((n, 80), dtype=torch.uint8, device="cuda")
            if n:
                from gravit_b200 import schedulers

                buf.copy_(schedulers.device_tensor(ptr, (n, 80), "|u1"))
            sends.append((counts, buf))
        mat = np.stack([c for c, _ in sends]).astype(np.int64)
        if mat.sum() == 0:
            break
        log_g.append(mat)
        for dst, e in enumerate(ranks):
            parts = []
            for src, (counts, buf) in enumerate(sends):
                off = int(counts[:dst].sum())
                parts.append(buf[off : off + int(counts[dst])])
            inbox = torch.cat(parts) if parts else None
            torch.cuda.synchronize()
            if inbox is not None and inbox.shape[0]:
                inbox = inbox.contiguous()
                e.domain_import(inbox.data_ptr(), inbox.shape[0])
    log_o = [m for m in log_o if m.sum() > 0]  # the oracle also logs the final all-zero round (termination test)
    assert len(log_g) == len(log_o) and len(log_g) >= 1
    for a, b in zip(log_g, log_o):
        assert np.array_equal(a, b)  # ray-routing counts, per round, per (src, dst)
    fb = np.zeros((d.height, d.width, 4), np.float32)
    for e in ranks:
        e.sync()
        fb += e.fb_read(d.width, d.height)
    fb[..., :3] = np.minimum(fb[..., :3], 1.0)
    assert np.allclose(fb, ref, atol=1e-6) and np.allclose(fb, ref1, atol=1e-6)
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) >= 50.0
    for e in ranks:
        e.close()


# ---- NVLink peer path: ranks simulated as several engine scenes in one process (attached by pointer, not by IPC) -------
def _connect(ranks):
    world = len(ranks)
    descs = b"".join(e.p2p_export(world) for e in ranks)
    for r, e in enumerate(ranks):
        e.p2p_attach(r, world, descs)


@pytest.mark.parametrize("world", [2, 3])
def test_image_scheduler_peer_gather_bit_identical(gpu, world):
    """Tracer<ImageScheduler> over the peer path: every rank renders its row tiles and stores them, clamped, into the
    root's framebuffer (replaces icetCompositeImage).  The root's image must equal the single-GPU frame bit for bit, for
    two consecutive frames (the go / done flags carry frame sequence numbers)."""
    from gravit_b200 import schedulers

    d = library.simple_scene(320, 200)
    ranks = [schedulers.build_domain_rank(d, 0, 1, flags=1)[0] for _ in range(world)]
    ref, _ = ranks[0].render(d.width, d.height)
    _connect(ranks)

    def frame(tile_rows, root):
        # Simulated ranks share ONE GPU and one host thread: the root's call is split (clear + go first, trace + wait after
        # the peers' calls) so that no spinning kernel waits for a launch the host has not made yet.
        ranks[root].image_frame_p2p(d.height, tile_rows, root, phases=1)
        for r, e in enumerate(ranks):
            if r != root:
                e.image_frame_p2p(d.height, tile_rows, root)
        ranks[root].image_frame_p2p(d.height, tile_rows, root, phases=2 | 4)
        for e in ranks:
            e.sync()
        return ranks[root].fb_read(d.width, d.height)

    for tile_rows in (3, 8, 8):  # consecutive frames: the go / done flags carry frame sequence numbers
        assert np.array_equal(frame(tile_rows, 0), ref)
    assert np.array_equal(frame(8, 1), ref)  # a non-zero root works the same way
    # parallel read-back (gvtb200_framebuffer_read_rows): every rank renders its row tiles and copies them to their place in ONE
    # host image; then every rank copies its share of the ROOT's gathered image (over the peer mapping) instead
    for tile_rows in (5, 8):
        host = np.zeros((d.height, d.width, 4), np.float32)
        for r, e in enumerate(ranks):
            e.render_tiles(d.height, tile_rows, r, world)
            e.fb_read_rows(host.ctypes.data, tile_rows, r, world)
        for e in ranks:
            e.sync()
        assert np.array_equal(host, ref)
    frame(8, 0)
    host = np.zeros((d.height, d.width, 4), np.float32)
    for r, e in enumerate(ranks):
        e.fb_read_rows(host.ctypes.data, 8, r, world, src_rank=0)
    for e in ranks:
        e.sync()
    assert np.array_equal(host, ref)
    for e in ranks:
        e.close()


@pytest.mark.parametrize("world", [2, 3])
def test_domain_scheduler_peer_exchange(gpu, world):
    """Tracer<DomainScheduler> over the peer path: the trace kernel stores migrating rays into the owner's inbound queue
    (replaces SendRays, DomainTracer.h:370-496).  Rays exchanged per round equal the oracle's routing counts, the
    composited image (touched pixels added into the root, then clamped) equals the single-process image."""
    from gravit_b200 import schedulers

    d = library.simple_scene(256, 256)
    flags = 1
    o = oracle_scene(d, flags)
    o.set_owner(np.arange(len(d.instances)) % world)
    ref, st = o.render(schedule=1, world_size=world, threads=8)
    per_round_o = [int(m.sum()) for m in o.route_log(world) if m.sum() > 0]
    ref1, _ = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(gpu, d, flags, world)
    _connect(ranks)
    for frame in range(2):  # twice: the inbound queues and flags must be reusable
        for e in ranks:
            e.domain_begin_frame()
        log = []
        for rnd in range(100):
            for e in ranks:
                e.domain_round_p2p(rnd)
            for e in ranks:
                e.sync()  # stands in for the all-reduce that ends a round
            log.append(sum(e.counters()["rays_exported"] for e in ranks))
            if schedulers.domain_terminated(log):
                break
        per_round = [b - a for a, b in zip([0] + log[:-1], log)][:-1]
        assert per_round == per_round_o and len(per_round) >= 1
        assert all(e.counters()["queue_overflows"] == 0 for e in ranks)
        for e in ranks[1:] + ranks[:1]:  # peers add their pixels, the root waits and clamps
            e.domain_composite_p2p(0)
        for e in ranks:
            e.sync()
        fb = ranks[0].fb_read(d.width, d.height)
        assert np.allclose(fb, ref, atol=1e-6) and np.allclose(fb, ref1, atol=1e-6)
        assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) >= 50.0
    for e in ranks:
        e.close()


def _persist_frame(ranks, root=0, tile_rows=4, grid=24):
    """One persistent DomainScheduler frame on simulated ranks: every rank's kernel must be resident at the same time on the
    one GPU they share (they wait for each other at the start barrier and exchange rays while running), so the grids are
    bounded; each scene launches on its own stream."""
    for e in ranks:
        e.domain_persist_config(grid_ctas=grid)
    for e in ranks:
        e.domain_frame_persist(root, tile_rows)
    for e in ranks:
        e.sync()


@pytest.mark.gpu
@pytest.mark.parametrize("world,scene", [(1, "simple"), (2, "simple"), (3, "simple"), (4, "gyroid"), (1, "bunny_obj")])
def test_domain_scheduler_persistent(gpu, world, scene):
    """The persistent DomainScheduler (gvtb200_domain_frame_persist; the reference's nearest relative is scheduler::AsyncDomain,
    tracer/Domain/DomainTracer.cpp:109-192): one kernel per rank and frame, sharded camera rays, rays stored into the serving
    rank's inbound queue and picked up by waiting lanes, shading in the same kernel, termination and composite on the device.
    The image in the root's framebuffer equals the single-process oracle image; hits, shadow rays and framebuffer adds equal
    the oracle's counts."""
    g = gpu
    g.check(g.lib().gvtb200_set_peer_timeout_ms(3000))
    d = {"simple": lambda: library.simple_scene(256, 256), "gyroid": lambda: library.gyroid_scene(48, 2, 320, 240),
         "bunny_obj": lambda: library.bunny_obj_scene(200, 160)}[scene]()
    flags = 1 if scene != "gyroid" else d.flags
    o = oracle_scene(d, flags)
    ref1, st = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(g, d, flags, world)
    if world > 1:
        _connect(ranks)
    for frame, root in enumerate([0, 0, world - 1]):  # three frames: counters, flags and queues are reusable; any root works
        _persist_frame(ranks, root)
        fb = ranks[root].fb_read(d.width, d.height)
        assert np.allclose(fb, ref1, atol=1e-6), (frame, float(np.abs(fb - ref1).max()))
        assert po.image_diff(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) == 0
        c = [e.counters() for e in ranks]
        for k in ("primary_hits", "shadow_generated", "fb_adds", "routed_initial", "camera_rays"):
            assert sum(x[k] for x in c) == st[k], (k, [x[k] for x in c], st[k])
        assert all(x["queue_overflows"] == 0 and x["stack_overflows"] == 0 for x in c)
        if world > 1:
            assert sum(x["rays_exported"] for x in c) > 0
    for e in ranks:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["replica", "mapped"])
def test_domain_scheduler_persistent_shared_blas(gpu, mode):
    """Several ranks serve one domain: the owner exports its BLAS, two other ranks import it -- one a replica in its own
    memory, or both a mapping of the owner's arrays -- and the server table splits the domain's rays between them by a hash of
    the ray.  Same image, same hit counts; the importing ranks really trace in the shared domain."""
    from gravit_b200 import schedulers

    g = gpu
    g.check(g.lib().gvtb200_set_peer_timeout_ms(3000))
    world = 3
    d = library.gyroid_scene(48, 2, 320, 240)
    o = oracle_scene(d, d.flags)
    ref1, st = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(g, d, d.flags, world)
    _connect(ranks)
    # a counted frame measures the load per instance (summed over the ranks), the planner spreads the hot ones
    for e in ranks:
        e.profile_counters(True)
    _persist_frame(ranks)
    load = sum(e.domain_persist_stats(len(d.instances))[2].astype(np.float64) for e in ranks)
    for e in ranks:
        e.profile_counters(False)
    assert load.sum() > 0
    tab_r, tab_w = schedulers.plan_servers(load, owner, world, max_replicas=2)
    assert (tab_r[:, 1] >= 0).any()
    before = schedulers.planned_rank_load(load, np.stack([owner, -np.ones(8), -np.ones(8), -np.ones(8)], 1).astype(np.int32), np.eye(4, dtype=np.float32)[[0] * 8], world)
    after = schedulers.planned_rank_load(load, tab_r, tab_w, world)
    assert after.max() < before.max()
    descs = [ranks[owner[k]]._keep[d.instances[k][0]].export() for k in range(len(d.instances))]
    keep = [schedulers.apply_server_plan(e, tab_r, tab_w, descs, r, mode=g.IMPORT_COPY if mode == "replica" else g.IMPORT_MAP) for r, e in enumerate(ranks)]
    assert any(keep)
    for frame in range(2):
        _persist_frame(ranks)
        fb = ranks[0].fb_read(d.width, d.height)
        assert np.allclose(fb, ref1, atol=1e-6)
        c = [e.counters() for e in ranks]
        for k in ("primary_hits", "shadow_generated", "fb_adds"):
            assert sum(x[k] for x in c) == st[k]
    # busy time and inbound-ray statistics exist for every rank
    for e in ranks:
        busy, rays, _ = e.domain_persist_stats()
        assert busy >= 0.0 and rays >= 0
    for e in ranks:
        e.close()
    del keep


@pytest.mark.gpu
@pytest.mark.parametrize("replicas", [1, 99])
def test_domain_scheduler_persistent_local_first(gpu, replicas):
    """Local-first replica plan (schedulers.plan_holders): every rank installs its OWN server table -- itself for the domains
    it holds, the holders by share for the others.  With a copy of everything everywhere no ray leaves its rank; with one
    replica per rank fewer do than with the owner alone.  Same image, same counters either way."""
    from gravit_b200 import schedulers

    g = gpu
    g.check(g.lib().gvtb200_set_peer_timeout_ms(3000))
    world = 3
    d = library.gyroid_scene(48, 2, 320, 240)
    o = oracle_scene(d, d.flags)
    ref1, st = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(g, d, d.flags, world)
    _connect(ranks)
    for e in ranks:
        e.profile_counters(True)
    _persist_frame(ranks)
    load = sum(e.domain_persist_stats(len(d.instances))[2].astype(np.float64) for e in ranks)
    exported0 = sum(e.counters()["rays_exported"] for e in ranks)
    for e in ranks:
        e.profile_counters(False)
    holders, share = schedulers.plan_holders(load, owner, world, max_replicas=replicas)
    descs = [ranks[owner[k]]._keep[d.instances[k][0]].export() for k in range(len(d.instances))]
    keep = [schedulers.apply_holder_plan(e, holders, share, descs, r, owner, mode=g.IMPORT_COPY) for r, e in enumerate(ranks)]
    assert any(keep)
    for frame in range(2):
        _persist_frame(ranks)
        fb = ranks[0].fb_read(d.width, d.height)
        assert np.allclose(fb, ref1, atol=1e-6)
        c = [e.counters() for e in ranks]
        for k in ("primary_hits", "shadow_generated", "fb_adds"):
            assert sum(x[k] for x in c) == st[k]
        exported = sum(x["rays_exported"] for x in c)
        if replicas == 99:
            assert all(len(h) == world for h in holders) and exported == 0
        else:
            assert 0 < exported < exported0
    for e in ranks:
        e.close()
    del keep


@pytest.mark.gpu
def test_domain_scheduler_persistent_with_bounces_and_lights(gpu):
    """rayMaxDepth 3, two lights, 3 simulated ranks: bounces continue in the lane, extra shadow rays go through the rank's own
    inbound queue, secondary rays migrate.  Random numbers are a function of (ray id, depth), not of who traces the ray, so the
    counters equal the oracle's and the image matches to fp32 reassociation."""
    g = gpu
    g.check(g.lib().gvtb200_set_peer_timeout_ms(3000))
    world = 3
    d = library.simple_scene(192, 192)
    d.depth = 3
    d.lights = d.lights + [((2.0, 2.0, 1.0), (0.6, 0.5, 0.4))]
    flags = 1
    o = oracle_scene(d, flags)
    ref1, st = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(g, d, flags, world)
    _connect(ranks)
    _persist_frame(ranks)
    c = [e.counters() for e in ranks]
    for k in ("primary_hits", "shadow_generated", "secondary_generated", "fb_adds"):
        assert abs(sum(x[k] for x in c) - st[k]) <= max(2, st[k] // 1000), (k, sum(x[k] for x in c), st[k])
    assert st["secondary_generated"] > 0
    fb = ranks[0].fb_read(d.width, d.height)
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) >= 50.0
    for e in ranks:
        e.close()


@pytest.mark.gpu
def test_domain_scheduler_with_bounces(gpu):
    """The round-synchronous DomainScheduler with rayMaxDepth 2 on 3 simulated ranks: secondary rays migrate as well (their
    random numbers are a function of ray id and depth, not of who traces them), so the number of rays that change rank equals
    the oracle's and the image matches the single-process image to fp32 reassociation."""
    from gravit_b200 import schedulers

    world = 3
    d = library.simple_scene(192, 192)
    d.depth = 2
    flags = 1
    o = oracle_scene(d, flags)
    o.set_owner(np.arange(len(d.instances)) % world)
    ref, st = o.render(schedule=1, world_size=world, threads=8)
    exchanged_o = int(sum(m.sum() for m in o.route_log(world)))
    ref1, _ = o.render(schedule=0, world_size=1, threads=8)
    ranks, owner = _standalone_ranks(gpu, d, flags, world)
    _connect(ranks)
    for e in ranks:
        e.domain_begin_frame()
    log = []
    for rnd in range(100):
        for e in ranks:
            e.domain_round_p2p(rnd)
        for e in ranks:
            e.sync()
        log.append(sum(e.counters()["rays_exported"] for e in ranks))
        if schedulers.domain_terminated(log):
            break
    assert sum(e.counters()["rays_exported"] for e in ranks) == exchanged_o and exchanged_o > 0
    assert sum(e.counters()["secondary_generated"] for e in ranks) == st["secondary_generated"] > 0
    assert all(e.counters()["queue_overflows"] == 0 for e in ranks)
    for e in ranks[1:] + ranks[:1]:
        e.domain_composite_p2p(0)
    for e in ranks:
        e.sync()
    fb = ranks[0].fb_read(d.width, d.height)
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) >= 50.0
    assert po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) >= 50.0
    for e in ranks:
        e.close()


@pytest.mark.gpu
def test_queue_overflow_is_an_error(gpu):
    """The reference never drops a ray (its queues are std::vectors, TracerBase.h:403-409).  With a queue budget too small
    for the frame (gvtb200_scene_set_queue_limit) the frame calls fail with GVTB200_ERR_CAPACITY -- from the wave path
    (hit / shadow queues) and from the persistent path (inbound queue) -- instead of returning an incomplete image, and the
    same scene renders correctly again once the budget is lifted."""
    g = gpu
    d = library.simple_scene(192, 192)
    d.depth = 2
    d.lights = d.lights + [((2.0, 2.0, 1.0), (0.6, 0.5, 0.4))]
    o = oracle_scene(d, 1)
    ref1, st = o.render(schedule=0, world_size=1, threads=8)
    ranks, _ = _standalone_ranks(g, d, 1, 1)
    e = ranks[0]
    fb, c = e.render(d.width, d.height)
    assert c["queue_overflows"] == 0 and po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) >= 50.0
    e.set_queue_limit(1000)
    with pytest.raises(g.EngineError, match="-4"):
        e.render(d.width, d.height)
    assert e.counters()["queue_overflows"] > 0
    e.render_async(d.height)
    with pytest.raises(g.EngineError, match="-4"):
        e.sync()
    with pytest.raises(g.EngineError, match="-4"):  # two lights + bounces: the extra rays go through the inbound queue
        e.domain_frame_persist(0, 4)
        e.sync()
    e.set_queue_limit(0)
    fb, c = e.render(d.width, d.height)
    assert c["queue_overflows"] == 0 and po.psnr(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref1)) >= 50.0
    e.domain_frame_persist(0, 4)
    e.sync()
    assert po.psnr(po.fb_to_ppm_bytes(e.fb_read(d.width, d.height)), po.fb_to_ppm_bytes(ref1)) >= 50.0
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("schedule,name", [(0, "Image"), (1, "Domain"), (3, "AsyncDomain")])
def test_cpp_multi_gpu_tracers_behind_api_render(gpu, schedule, name):
    """Renderer.cpp:71-96: api::render picks Tracer<ImageScheduler> / Tracer<DomainScheduler> from the renderer's schedule
    type.  With api::setGpus(n) the C++ host layer drives n GPUs of ONE process (one host thread per GPU, peer access
    instead of IPC, no torch, no MPI): the 25-instance golden scene rendered on 2 (and, if present, 4) GPUs equals the oracle's
    single-process image, for every schedule type.  Needs >= 2 GPUs (skipped on a single-GPU box)."""
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    g = gpu
    d = library.simple_scene(256, 256)
    o = oracle_scene(d, 1)
    ref, st = o.render(schedule=0, world_size=1, threads=8)
    s = library.apply(d, g.GvtScene(flags=1))
    one = s.render()
    assert np.array_equal(one, ref)
    s.set_schedule(schedule)
    try:
        for n in sorted({2, min(4, n_dev)}):
            s.set_gpus(n)
            for frame in range(2):
                fb = s.render()
                assert np.allclose(fb, ref, atol=1e-6), (n, frame, float(np.abs(fb - ref).max()))
                assert po.image_diff(po.fb_to_ppm_bytes(fb), po.fb_to_ppm_bytes(ref)) == 0
    finally:
        s.set_gpus(1)
    assert np.array_equal(s.render(), ref)  # and back on one GPU


@pytest.mark.gpu
def test_conf_file_scene_end_to_end(gpu, tmp_path):
    """a data/*.conf style scene file (OBJ + MTL, two instances of one mesh, point light) loaded by the C++
    ConfigFileLoader / ObjReader, rendered through api::render, equals the oracle fed with the same arrays; writeimage
    round-trips through the PPM writer."""
    import math

    g = gpu
    fx = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "readers")
    A = g.api()
    g.acheck(A.gvtapi_reset())
    g.acheck(A.gvtapi_set_engine_flags(0, 0))
    info = g.load_conf(os.path.join(fx, "scene.conf"))
    W, H = info["width"], info["height"]
    g.acheck(A.gvtapi_render(b"Renderer"))
    fb = np.zeros((H, W, 4), np.float32)
    g.acheck(A.gvtapi_framebuffer(g._fp(fb), fb.size, None, None))

    m = g.read_obj(os.path.join(fx, "box.obj"))
    mats = np.zeros((len(m["materials"]), 8), np.float32)
    mats[:, 0], mats[:, 1:4], mats[:, 4:7], mats[:, 7] = m["materials"][:, 9], m["materials"][:, 3:6], m["materials"][:, 6:9], 1.0
    o = po.OracleScene(flags=0)
    mid = o.add_mesh(m["vertices"], m["faces"], one_based=False)
    o.mesh_face_materials(mid, mats, m["face_mat"])
    o.add_instance(mid, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, np.float32(-0.8), 0, 0, 1])
    o.add_instance(mid, [0.5, 0, 0, 0, 0, 1.5, 0, 0, 0, 0, 0.5, 0, np.float32(0.8), 0, 0, 1])
    o.add_point_light((0.5, 2.5, 3.0), (1.0, 1.0, 1.0))
    o.set_camera((0.2, 0.6, 3.5), (0.0, 0.2, 0.0), (0, 1, 0), float(np.float32(40.0 * math.pi / 180.0)), 1, 1, 0.0)
    o.set_film(W, H)
    o.commit()
    ref, st = o.render(threads=4)
    assert st["primary_hits"] > 500 and (ref[..., :3].reshape(-1, 3).max(0) > 0).all()
    assert np.array_equal(fb, ref)
    out = str(tmp_path / "conf_scene")
    g.acheck(A.gvtapi_writeimage(b"Renderer", out.encode()))
    w, h, pix = po.read_ppm(open(out + ".ppm", "rb").read())
    assert (w, h) == (W, H) and np.array_equal(pix, po.fb_to_ppm_bytes(ref))
    g.acheck(A.gvtapi_reset())
