"""The oracle is pinned against the reference's own fixtures: the two golden images
(Test/CTESTtest/data/{simple,bunny}.ppm, tolerance 300, CMakeLists.txt:666) and the reference pieces that compile
standalone (oracle/_ref: Shade, Box3D, RayPacketIntersection<8>, Mesh, GLM)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from scenes import library
from tests.scene_util import golden_ppm, oracle_scene


def _render_bytes(desc, flags, threads=4):
    s = oracle_scene(desc, flags)
    fb, st = s.render(threads=threads)
    return po.fb_to_ppm_bytes(fb), st, s


def test_known_sizes(built):
    assert po.lib().gvto_sizeof_ray() == 80  # sizeof(gvt::render::actor::Ray)
    assert po.RAY_DTYPE.itemsize == 80


def test_golden_simple_smooth(built):
    w, h, gold = golden_ppm("simple")
    img, st, s = _render_bytes(library.simple_scene(), po.FLAG_SMOOTH)
    assert (w, h) == (512, 512)
    assert po.image_diff(img, gold) < 300  # the reference's CTest tolerance
    assert po.psnr(img, gold) > 50.0
    assert st["adapter_calls"] == 37


def test_golden_bunny_smooth(built):
    w, h, gold = golden_ppm("bunny")
    img, st, s = _render_bytes(library.bunny_obj_scene(), po.FLAG_SMOOTH)
    assert po.image_diff(img, gold) < 300
    assert po.psnr(img, gold) > 50.0
    assert st["routed_initial"] == 155232


def test_flat_mode_differs_from_golden(built):
    """HEAD hard-codes FLAT_SHADING (EmbreeMeshAdapter.cpp:75); the goldens pre-date it."""
    _, _, gold = golden_ppm("bunny")
    img, _, _ = _render_bytes(library.bunny_obj_scene(), 0)
    assert 30.0 < po.psnr(img, gold) < 40.0


def test_bvh_equals_brute_force(built):
    d = library.bunny_obj_scene(64, 64)
    s = oracle_scene(d)
    rays = s.camera_rays()
    a = s.intersect(0, rays, brute=False)
    b = s.intersect(0, rays, brute=True)
    assert np.array_equal(a, b)


def test_packet_routing_equals_single_ray(built):
    s = oracle_scene(library.simple_scene(128, 128))
    rays = s.camera_rays()
    n8, t8 = s.route(rays, -1, simd=8)
    n1, t1 = s.route(rays, -1, simd=1)
    assert np.array_equal(n8, n1) and np.array_equal(t8, t1)


def test_domain_schedule_equals_image_schedule(built):
    """the P-rank DomainScheduler image equals the single-process image."""
    d = library.simple_scene(128, 128)
    s = oracle_scene(d, po.FLAG_SMOOTH)
    fb1, st1 = s.render(schedule=0, world_size=1, threads=2)
    fb2, st2 = s.render(schedule=1, world_size=2, threads=2)
    assert np.allclose(fb1, fb2, atol=1e-6)
    assert st2["rays_sent"] > 0


ref = po.ref_lib() if True else None


@pytest.mark.skipif(po.ref_lib() is None, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_vs_compiled_reference_pieces(built):
    import ctypes as C

    R = po.ref_lib()
    assert R.gvtref_sizeof_ray() == 80 and R.gvtref_sizeof_box() == 32 and R.gvtref_vec3_length() == 3
    rng = np.random.default_rng(0)
    # Shade(): Material.cpp:90-139 with a PointLight
    for mtype in (0, 1, 2):
        for _ in range(200):
            ray = np.zeros(1, po.RAY_DTYPE)
            ray["origin"], ray["direction"] = rng.normal(size=3), rng.normal(size=3)
            ray["direction"] /= np.linalg.norm(ray["direction"])
            ray["t"], ray["w"] = rng.random() * 3, rng.random()
            N = rng.normal(size=3).astype(np.float32)
            N /= np.linalg.norm(N)
            kd, ks = rng.random(3).astype(np.float32), rng.random(3).astype(np.float32)
            lp, lc = (rng.normal(size=3) * 2).astype(np.float32), rng.random(3).astype(np.float32)
            alpha = float(np.float32(rng.random() * 20))
            o1, o2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
            r1 = po.lib().gvto_shade(mtype, po._fp(kd), po._fp(ks), alpha, po._fp(ray), po._fp(N), 0, po._fp(lp), po._fp(lc), po._fp(o1))
            r2 = R.gvtref_shade(mtype, po._fp(kd), po._fp(ks), alpha, po._fp(ray), po._fp(N), po._fp(lp), po._fp(lc), po._fp(o2))
            assert bool(r1) == bool(r2)
            if r1:
                if mtype == 0:
                    assert np.array_equal(o1, o2)
                else:
                    assert np.allclose(o1, o2, rtol=1e-5, atol=1e-7)


@pytest.mark.skipif(po.ref_lib() is None, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_routing_vs_reference_packet_box_test(built):
    """RayPacketIntersection<8>::intersect (actor/RayPacket.h:111-211) of the compiled reference against the oracle's routing of
    a one-instance scene whose instance box is the same box: same hit lanes, same entry t (update = true: the lane's t
    becomes the entry distance, :202), for rays from outside, from inside the box, parallel to faces and with zero
    components."""
    import ctypes as C

    R = po.ref_lib()
    R.gvtref_packet_box.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(3)
    lo, hi = np.array([-0.5, -0.25, 0.1], np.float32), np.array([0.75, 0.5, 0.9], np.float32)
    verts = np.stack([lo, hi, np.array([lo[0], hi[1], lo[2]], np.float32)])
    s = po.OracleScene(flags=0)
    s.add_mesh(verts, np.array([[1, 2, 3]], np.uint32), one_based=True)
    s.add_instance(0, np.eye(4, dtype=np.float32).T.reshape(16))
    s.add_point_light((0, 0, 5), (1, 1, 1))
    s.set_camera((0, 0, 3), (0, 0, 0), (0, 1, 0), 0.8, 1, 1, 0.0)
    s.set_film(8, 8)
    s.commit()
    box6 = np.concatenate([lo, hi]).astype(np.float32)
    n_checked = n_hit = 0
    for trial in range(300):
        rays = np.zeros(8, po.RAY_DTYPE)
        o = rng.uniform(-2, 2, (8, 3)).astype(np.float32)
        if trial % 3 == 0:
            o[:4] = rng.uniform(lo, hi, (4, 3)).astype(np.float32)  # origins inside the box: never entered (Q6)
        d = rng.normal(size=(8, 3)).astype(np.float32)
        if trial % 5 == 0:
            d[:, rng.integers(0, 3)] = 0.0  # axis-parallel: 1/0 = inf in the slab test
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays["origin"], rays["direction"] = o, d
        rays["t_min"], rays["t_max"], rays["t"] = 1e-6, np.float32(3.4028235e38), np.float32(3.4028235e38)
        if trial % 7 == 0:
            rays["t_max"][:3] = 3.0  # shadow rays carry t_max = 3 (Q2): boxes further away are not entered
        hit, t = np.zeros(8, np.int32), np.zeros(8, np.float32)
        R.gvtref_packet_box(po._fp(rays), 8, po._fp(box6), 1, po._fp(hit), po._fp(t))
        nxt, to = s.route(rays, -1)
        assert np.array_equal(hit == 1, nxt == 0) and set(np.unique(hit)) <= {1, -1}, (trial, hit, nxt)  # :194: 1 = hit, -1 = miss
        assert np.array_equal(t[hit == 1], to[nxt == 0])
        n_checked += 8
        n_hit += int((hit == 1).sum())
    assert 0 < n_hit < n_checked


@pytest.mark.skipif(po.ref_lib() is None, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_mesh_and_instance_vs_reference(built):
    """Mesh::addFace's degenerate-face filter, generateNormals and computeBoundingBox (data/primitives/Mesh.cpp:102-154,186-199)
    and api::addInstance's arithmetic (api/api.cpp:300-321: glm::inverse, transpose(inverse(mat3)), two-corner instance box,
    centroid) of the compiled reference against the oracle, bit for bit."""
    import ctypes as C

    R = po.ref_lib()
    rng = np.random.default_rng(11)
    for trial in range(20):
        nv, nt = int(rng.integers(4, 60)), int(rng.integers(1, 120))
        verts = rng.normal(size=(nv, 3)).astype(np.float32)
        verts[rng.integers(0, nv)] = verts[0]  # two vertices at one position: faces joining them are dropped
        tris = rng.integers(1, nv + 1, (nt, 3)).astype(np.uint32)
        tris[::7, 1] = tris[::7, 0]  # repeated index
        f_ref, n_ref, b_ref = np.zeros((nt, 3), np.int32), np.zeros((nv, 3), np.float32), np.zeros(6, np.float32)
        nf = R.gvtref_mesh(po._fp(verts), nv, po._fp(tris), nt, po._fp(f_ref), po._fp(n_ref), po._fp(b_ref))
        s = po.OracleScene(flags=0)
        s.add_mesh(verts, tris, one_based=True)
        f_o, n_o, b_o = s.mesh_get(0, nv)
        assert nf == len(f_o) <= nt and np.array_equal(f_ref[:nf], f_o)
        assert np.array_equal(n_ref.view(np.uint32), n_o.view(np.uint32)) or np.array_equal(np.isnan(n_ref), np.isnan(n_o)) and np.array_equal(n_ref[~np.isnan(n_ref)], n_o[~np.isnan(n_o)])
        assert np.array_equal(b_ref, b_o)
        # an affine instance matrix: rotation * non-uniform scale + translation (column-major, glm::make_mat4)
        a = rng.normal(size=(3, 3)).astype(np.float32)
        q, _ = np.linalg.qr(a.astype(np.float64))
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = (q * rng.uniform(0.3, 2.5, 3)).astype(np.float32)
        m[:3, 3] = rng.uniform(-3, 3, 3).astype(np.float32)
        m16 = np.ascontiguousarray(m.T).reshape(16)
        minv_r, normi_r, box_r, cen_r = np.zeros(16, np.float32), np.zeros(9, np.float32), np.zeros(6, np.float32), np.zeros(3, np.float32)
        R.gvtref_instance(po._fp(m16), po._fp(b_ref), po._fp(minv_r), po._fp(normi_r), po._fp(box_r), po._fp(cen_r))
        s.add_instance(0, m16)
        _, minv_o, normi_o, box_o = s.instance_get(0)
        assert np.array_equal(minv_r, minv_o) and np.array_equal(normi_r, normi_o) and np.array_equal(box_r, box_o)
        # normalize(normi * n) as applied to normals (EmbreeMeshAdapter.cpp:504,518)
        R.gvtref_normal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        po.lib().gvto_normal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for _ in range(10):
            v = rng.normal(size=3).astype(np.float32)
            o1, o2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
            R.gvtref_normal(po._fp(normi_r), po._fp(v), po._fp(o1))
            po.lib().gvto_normal(po._fp(normi_o), po._fp(v), po._fp(o2))
            assert np.array_equal(o1, o2)


@pytest.mark.skipif(po.ref_lib() is None, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_box_ops_vs_reference(built):
    """Box3D::merge, surfaceArea and wideRangingBoxDir (data/primitives/BBox.cpp:55-63,118-133) -- what the top-level BVH build
    (data/accel/BVH.cpp:77-216) is made of -- of the compiled reference against the oracle's Box, including point boxes and
    boxes whose extent overflows."""
    import ctypes as C

    R = po.ref_lib()
    for L in (R, po.lib()):
        fn = L.gvtref_box_merge_area if L is R else L.gvto_box_merge_area
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        fn.restype = C.c_float
    rng = np.random.default_rng(5)
    FM = np.float32(3.4028235e38)
    for trial in range(400):
        a = np.sort(rng.normal(size=(2, 3)).astype(np.float32), axis=0).reshape(6)
        b = np.sort((rng.normal(size=(2, 3)) * rng.uniform(0.01, 5)).astype(np.float32), axis=0).reshape(6)
        if trial % 10 == 0:
            a = np.array([-FM, -FM, -FM, FM, FM, FM], np.float32)  # the largest box (the two-corner constructor sorts its corners,
            #                                                        BBox.cpp:41-46, so the empty accumulator cannot be passed in)
        if trial % 13 == 0:
            b[3:] = b[:3]  # a degenerate (point) box
        o1, o2, w1, w2 = np.zeros(6, np.float32), np.zeros(6, np.float32), C.c_int(-1), C.c_int(-2)
        s1 = R.gvtref_box_merge_area(po._fp(a), po._fp(b), po._fp(o1), C.byref(w1))
        s2 = po.lib().gvto_box_merge_area(po._fp(a), po._fp(b), po._fp(o2), C.byref(w2))
        assert np.array_equal(o1, o2) and w1.value == w2.value
        assert np.float32(s1) == np.float32(s2) or (np.isnan(s1) and np.isnan(s2))
