"""Pins the oracle (oracle/oracle.cpp) with every known answer that can be derived from the reference's source
 and with the published vectors of the two RNG algorithms it restates.  The reference ships
no tests or golden vectors of its own, and cannot be built here; third-party arithmetic stays "parity unpinned"."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from bindings import (RNG_CHACHA, RNG_PHILOX, OracleScene, oracle_get_ray, oracle_lib, oracle_philox, oracle_resolve,
                      oracle_scatter)


def test_ka1_miss_pixel_is_b4b4b4(pt):
    # background GRAY 0.5 (renderer.rs:61) -> sqrt -> *255 -> truncate (renderer.rs:115-118, color.rs:87-93)
    img = np.full((1, 3), 0.5, np.float32)
    assert oracle_resolve(img)[0] == 0x00B4B4B4
    # the same through a render of an empty scene
    s = pt.Scene()
    s.set_camera((0, 0, 0), (0, 0, -1), (0, 1, 0), 60.0, 1.0)
    out, st = OracleScene(s).render(s.camera, 4, 4, 2, 5, rng_mode=RNG_CHACHA)
    assert (out == 0.5).all() and st.rays == 4 * 4 * 2
    assert (oracle_resolve(out) == 0x00B4B4B4).all()


def test_resolve_clamp_nan_and_truncation():
    img = np.array([[4.0, -1.0, np.nan], [1.0, 0.25, 0.0], [0.999, 1e-9, 2.0]], np.float32)
    out = oracle_resolve(img)
    assert out[0] == (255 << 16) | (0 << 8) | 0  # sqrt(-1) = NaN -> 0, NaN -> 0
    assert out[1] == (255 << 16) | (127 << 8) | 0  # sqrt(.25)*255 = 127.5 -> 127
    assert out[2] == (int(math.sqrt(np.float32(0.999)) * 255) << 16) | (0 << 8) | 255


@pytest.mark.parametrize("name,tris,nodes,leaves,depth", [("RayTracingText.obj", 4748, 3351, 1676, 11),
                                                          ("teapot/teapot.obj", 6320, 4095, 2048, 11)])
def test_ka2_reference_bvh_shape(pt, scenes_dir, name, tris, nodes, leaves, depth):
    s = pt.Scene()
    m = s.add_material(pt.lambertian((0.5, 0.5, 0.5)))
    s.add_obj(os.path.join(scenes_dir, name), m)
    assert len(s.mesh(0)) == tris  # no triangle fails the degenerate filter (mesh_object.rs:128-134)
    n, l, d, dead, order = OracleScene(s).mesh_bvh_info(0, tris)
    assert (n, l, d) == (nodes, leaves, depth)
    assert sorted(order.tolist()) == list(range(tris))


def test_flat_node_culling_counts(pt, scenes_dir):
    # SURVEY finding 3: 1,982 of 4,748 triangles of the default scene's mesh sit under a zero-extent node
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "semesterbild.json"))
    _, _, _, dead, _ = OracleScene(s).mesh_bvh_info(3, 4748)
    assert int(dead.sum()) == 1982


def test_ka3_sphere_light_radiance(pt, scenes_dir):
    # power / (4 pi^2 r^2), parser.rs:566-575 (pi squared, as written)
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "veach-mis", "scene.json"))
    mats = s.materials
    got = sorted(round(mats[o.material].albedo[0], 3) for o in s.objects if o.type == pt.OBJ_SPHERE)
    want = sorted(round(300.0 / (4 * math.pi ** 2 * r * r), 3) for r in (1.0, 0.5, 0.05))
    assert got == pytest.approx(want, rel=1e-4)
    assert all(mats[o.material].type == pt.MAT_EMISSIVE for o in s.objects if o.type == pt.OBJ_SPHERE)


def test_ka4_camera_centre_ray_is_forward(pt):
    cam = pt.camera_new((1, 2, 3), (4, 1, -2), (0, 1, 0), 40.0, 1.5)
    o, d = oracle_get_ray(cam, 0.5, 0.5)
    assert np.allclose(d, list(cam.forward), atol=1e-7) and np.allclose(o, [1, 2, 3])
    # row 0 is the top: v = 0 looks up
    _, d_top = oracle_get_ray(cam, 0.5, 0.0)
    assert float(np.dot(d_top, list(cam.true_up))) > 0


@pytest.mark.parametrize("mat", ["lambert", "dielectric"])
def test_ka5_furnace_energy_conserving_materials(pt, mat):
    # a unit sphere alone under the constant 0.5 background must look like the background (material.rs:69,129)
    s = pt.Scene()
    m = s.add_material(pt.lambertian((1, 1, 1)) if mat == "lambert" else pt.dielectric(1.5))
    s.add_sphere((0, 0, 0), 1.0, m)
    s.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 30.0, 1.0)
    img, _ = OracleScene(s).render(s.camera, 32, 32, 16, 64, rng_mode=RNG_CHACHA)
    assert img.min() > 0.499 and img.max() < 0.501


def test_ka6_schlick_r0(pt):
    # Dielectric at normal incidence, ior 1.5: reflect iff 0.04 > u0 (material.rs:143-145,221-227)
    m = pt.dielectric(1.5)
    d, n, p = np.array([[0, 0, -1.0]]), np.array([[0, 0, 1.0]]), np.zeros((1, 3))
    refl = oracle_scatter(m, d, p, n, [1], [[0.039, 0, 0, 0]])
    refr = oracle_scatter(m, d, p, n, [1], [[0.041, 0, 0, 0]])
    assert refl[2][0][2] > 0.99 and refr[2][0][2] < -0.99
    assert (refl[3] == 1).all()


def test_chacha20_rfc7539_block_vector():
    # RFC 7539 §2.3.2: key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00, 20 rounds
    key = np.frombuffer(bytes(range(32)), dtype="<u4").copy()
    counter = 1 | (0x09000000 << 32)
    stream = 0x4A000000
    out = np.zeros(16, np.uint32)
    oracle_lib().orc_chacha_block(key.ctypes.data, C.c_uint64(counter), C.c_uint64(stream), 20, out.ctypes.data)
    want = [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3, 0xc7f4d1c7, 0x0368c033, 0x9aaa2204, 0x4e6cd4c3,
            0x466482d2, 0x09aa9f07, 0x05d7c214, 0xa2028bd9, 0xd19c12b5, 0xb94e16de, 0xe883d0cb, 0x4e3c50a2]
    assert out.tolist() == want


def test_philox4x32_10_random123_vectors():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert oracle_philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle_philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle_philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_chacha_std_rng_stream_properties():
    # StdRng::seed_from_u64 (PCG32 key expansion) + ChaCha12: deterministic per seed, rows independent
    L = oracle_lib()
    a, b, c = (np.zeros(256, np.uint32) for _ in range(3))
    L.orc_chacha_stream(5, a.ctypes.data, 256)
    L.orc_chacha_stream(5, b.ctypes.data, 256)
    L.orc_chacha_stream(6, c.ctypes.data, 256)
    assert (a == b).all() and (a != c).mean() > 0.99
    u = (a >> 8).astype(np.float64) / (1 << 24)
    assert 0.4 < u.mean() < 0.6


def test_oracle_rng_modes_agree_statistically(pt, scenes_dir):
    # Philox (the GPU's stream, direct sphere sampling) and ChaCha (the reference's stream, rejection sampling) must
    # describe the same estimator
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "cornell-box", "scene.json"))
    orc = OracleScene(s)
    a, sa = orc.render(s.camera, 48, 48, 64, 8, rng_mode=RNG_CHACHA)
    b, sb = orc.render(s.camera, 48, 48, 64, 8, rng_mode=RNG_PHILOX, seed=3)
    assert abs(a.mean() - b.mean()) / a.mean() < 0.02
    assert abs(sa.rays - sb.rays) / sa.rays < 0.01


def test_oracle_sample_ranges_partition_the_render(pt, scenes_dir):
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "cornell-box", "scene.json"))
    orc = OracleScene(s)
    full, st = orc.render(s.camera, 24, 24, 8, 6, seed=9)
    lo, _ = orc.render(s.camera, 24, 24, 8, 6, seed=9, sample_begin=0, sample_end=3)
    hi, _ = orc.render(s.camera, 24, 24, 8, 6, seed=9, sample_begin=3, sample_end=8)
    assert np.allclose(lo + hi, full, rtol=1e-5, atol=1e-6)
