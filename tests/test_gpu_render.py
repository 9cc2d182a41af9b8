"""The wavefront renderer on a B200 against the oracle: same-stream image parity, statistical parity against the
reference's own RNG stream, furnace tests per material, sharding invariance, edge cases, the host-facing entry points.

Tolerances (stated here, as the north star asks):
  same Philox stream   : >= 99.5 % of pixels within rtol 1e-3 / atol 1e-4 of the oracle, ray counts within 0.1 %
                         (the two sides differ only by ulp-level libm differences that can flip a rare branch)
  reference ChaCha run : relMSE(GPU_spp, R) <= 1.3 * relMSE(oracle_spp, R) with R = oracle at 16x spp,
                         relMSE(a,b) = mean((a-b)^2 / (b^2 + 1e-2)); per-channel mean within 1 %; rays/path within 1 %
  furnace              : sphere-pixel mean within 0.5 % of the oracle's (exactly 0.5 +- 1e-3 for Lambertian albedo 1
                         and Dielectric)
"""
import ctypes as C
import os

import numpy as np
import pytest

from bindings import compare_hits, RNG_CHACHA, RNG_PHILOX, OracleScene, oracle_resolve

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "scenes")


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def _same_stream(pt, scene, w, h, spp, depth, seed=3):
    cs = scene.to_core().commit(0)
    st = scene.render_settings(width=w, height=h, spp=spp, max_depth=depth, seed=seed)
    img, stats = cs.render(scene.camera, st)
    ref, ostats = OracleScene(scene).render(scene.camera, w, h, spp, depth, rng_mode=RNG_PHILOX, seed=seed)
    close = np.isclose(img, ref, rtol=1e-3, atol=1e-4).all(axis=2).mean()
    assert close >= 0.995, close
    assert stats.paths == w * h * spp == ostats.paths
    assert abs(int(stats.rays) - int(ostats.rays)) <= 1e-3 * ostats.rays
    assert abs(img.mean() - ref.mean()) <= 2e-3 * ref.mean()
    return img, ref, stats


@pytest.mark.parametrize("name,w,h,spp,depth", [("cornell-box/scene.json", 128, 128, 16, 8),     # C1 shrunk
                                                ("semesterbild.json", 200, 150, 8, 30),            # C2 shrunk
                                                ("veach-mis/scene.json", 160, 90, 16, 16)])        # C4 shrunk
def test_same_stream_image_parity(pt, name, w, h, spp, depth):
    _same_stream(pt, pt.load_scene_from_json(os.path.join(SCENES, name)), w, h, spp, depth)


def test_shipped_teapot_scene_with_loader_extensions(pt):
    # the reference's own tungsten/teapot scene (infinite_sphere -> sky, WO3 meshes with 16-byte records: 124,840
    # triangles, plastic + checker, HDR sky): closest hits and the same-stream image against the oracle
    s = pt.load_scene_from_json(os.path.join(SCENES, "teapot", "scene.json"), pt.LOAD_INFINITE_SPHERE_SKY | pt.LOAD_WO3_STRIDE16)
    cs = s.to_core().commit(0)
    st = s.render_settings(width=320, height=180, spp=1, max_depth=1, seed=3)
    o, d = cs.primary_rays(s.camera, st, 0)
    got, _ = cs.intersect(o, d)
    want = OracleScene(s).intersect(pt, o, d)
    r = compare_hits(got, want)
    assert r["id_mismatch"] == 0 and r["t_mismatch"] == 0 and r["bit_exact_records"] == r["both_hit"] > 20000, r
    _same_stream(pt, s, 160, 90, 8, 12, seed=3)


def test_config_c3_derived_same_stream(pt):
    # BASELINE config C3 as derived in SURVEY.md 8d (teapot.obj x30, GGX copper, checker quad), shrunk frame
    from raytracer_rust_b200 import workloads
    _, s = workloads.workload("C3")
    img, ref, stats = _same_stream(pt, s, 192, 108, 8, 16, seed=5)
    assert stats.rays > 1.5 * stats.paths  # the copper teapot and the floor do scatter


def test_config_c5_same_stream(pt):
    # BASELINE config C5's generator (height field + glass sphere + GGX-Al cube + emissive quad), 300x300 cells =
    # 180,000 triangles, shrunk frame, the config's own depth 16
    s = pt.synthetic_scene(cells=300)
    img, ref, stats = _same_stream(pt, s, 192, 108, 4, 16, seed=6)
    assert stats.rays > 1.5 * stats.paths


def test_work_stealing_does_not_change_the_result(pt, monkeypatch):
    # k_traverse splits the last traversals of a launch over the idle lanes of their warp (in-warp work stealing); the
    # closest hit is the minimum of a 64-bit (t, DFS position) key whatever the split, so image, ray count and hit records
    # must be those of the unsplit walk
    s = pt.load_scene_from_json(os.path.join(SCENES, "semesterbild.json"))
    cs = s.to_core().commit(0)
    base = dict(width=200, height=150, spp=8, max_depth=30, seed=12)
    a, sa = cs.render(s.camera, s.render_settings(**base))
    st = s.render_settings(width=320, height=240, spp=1, max_depth=1, seed=1)
    o, d = cs.primary_rays(s.camera, st, 0)
    ha, _ = cs.intersect(o, d)
    monkeypatch.setenv("PTC_STEAL", "0")
    b, sb = cs.render(s.camera, s.render_settings(**base))
    hb, _ = cs.intersect(o, d)
    assert sa.rays == sb.rays and sa.paths == sb.paths == 200 * 150 * 8 and sa.iterations == sb.iterations
    assert np.array_equal(a, b)
    assert ha.tobytes() == hb.tobytes()


def test_segment_ordering_does_not_change_the_result(pt, monkeypatch):
    # the launches of a render are ordered per SEGMENT (a block waits for its own segment's flag), and a
    # traversal launch may hand (segment, part) items to whichever block is free; PTC_DATAFLOW=0 goes back to whole-launch
    # waits, PTC_TRAV_PARTS forces / forbids the split.  Paths are keyed by (pixel, sample) and the film is fixed-point, so
    # image, ray count and iteration count must be the same in all four combinations — with the whole job in flight and
    # with the smallest pool (smaller than the job: segments are topped up over several iterations, the split path runs
    # until the supply is exhausted)
    s = pt.load_scene_from_json(os.path.join(SCENES, "semesterbild.json"))
    cs = s.to_core().commit(0)
    for extra in ({}, {"pool_paths": 592 * 256}):
        kw = dict(width=200, height=150, spp=8, max_depth=30, seed=12, **extra)
        ref, sref = cs.render(s.camera, s.render_settings(**kw))
        assert sref.paths == 200 * 150 * 8
        for dataflow, parts in (("0", "1"), ("1", "4"), ("0", "4"), ("1", "1")):
            monkeypatch.setenv("PTC_DATAFLOW", dataflow)
            monkeypatch.setenv("PTC_TRAV_PARTS", parts)
            img, st = cs.render(s.camera, s.render_settings(**kw))
            assert (st.rays, st.paths, st.iterations) == (sref.rays, sref.paths, sref.iterations)
            assert np.array_equal(img, ref)
        monkeypatch.delenv("PTC_DATAFLOW")
        monkeypatch.delenv("PTC_TRAV_PARTS")


def test_render_is_ordered_on_the_callers_stream(pt):
    # ptc_render_accumulate with the NULL (legacy default) stream: work queued there beforehand — a long kernel, then the
    # zeroing of the film — must be ordered before the render's film updates (the render used to run on a
    # private non-blocking stream, so the memset could land in the middle of it)
    torch = pytest.importorskip("torch")
    s = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    cs = s.to_core().commit(0)
    w = h = 96
    st = s.render_settings(width=w, height=h, spp=4, max_depth=5, seed=8)
    want, _ = cs.render(s.camera, st)
    accum = torch.full((w * h * 3,), 1000.0, dtype=torch.float32, device="cuda:0")
    for stream_ptr in (None, torch.cuda.current_stream().cuda_stream):
        accum.fill_(1000.0)
        torch.cuda.synchronize()
        torch.cuda._sleep(400_000_000)  # ~0.2 s on the default stream
        accum.zero_()
        cs.render_accumulate(s.camera, st, accum.data_ptr(), stream_ptr)
        got = accum.cpu().numpy().reshape(h, w, 3) / 4.0
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6), float(np.abs(got - want).max())


@pytest.mark.parametrize("name,spp", [("C2", 4), ("C3", 2), ("C4", 2), ("C5", 1)])
def test_native_frame_same_stream(pt, name, spp):
    # BASELINE configs C2 (800x600, depth 30), C3 (derived teapot, 1280x720, depth 64), C4 (1280x720, depth 16) and C5 (the
    # full 2,000,000-triangle mesh, 3840x2160, depth 16) at their NATIVE frame and bounce limit — only the sample count is
    # reduced, to what the CPU oracle renders in a few seconds
    from raytracer_rust_b200 import workloads
    _, s = workloads.workload(name)
    w, h, _, depth = s.settings
    _same_stream(pt, s, w, h, spp, depth, seed=2)


def test_config_c1_full_size_same_stream(pt):
    # BASELINE config C1 exactly: Cornell box 256x256, 16 spp, 8 bounces
    s = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    img, ref, stats = _same_stream(pt, s, 256, 256, 16, 8, seed=0)
    assert stats.paths == 1048576


@pytest.mark.parametrize("name,w,h,spp,depth", [("cornell-box/scene.json", 64, 64, 16, 8), ("semesterbild.json", 160, 120, 8, 30)])
def test_statistical_parity_with_reference_stream(pt, name, w, h, spp, depth):
    s = pt.load_scene_from_json(os.path.join(SCENES, name))
    orc = OracleScene(s)
    R, _ = orc.render(s.camera, w, h, spp * 16, depth, rng_mode=RNG_CHACHA)
    o, ostats = orc.render(s.camera, w, h, spp, depth, rng_mode=RNG_CHACHA)
    cs = s.to_core().commit(0)
    g, gstats = cs.render(s.camera, s.render_settings(width=w, height=h, spp=spp, max_depth=depth, seed=11))
    assert relmse(g, R) <= 1.3 * relmse(o, R)
    for c in range(3):
        assert abs(g[..., c].mean() - R[..., c].mean()) <= 0.01 * R[..., c].mean()
    assert abs(gstats.rays / gstats.paths - ostats.rays / ostats.paths) <= 0.01 * ostats.rays / ostats.paths


def _furnace_scene(pt, material):
    s = pt.Scene()
    m = s.add_material(material)
    s.add_sphere((0, 0, 0), 1.0, m)
    s.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 30.0, 1.0)
    return s


@pytest.mark.parametrize("which", ["lambert", "dielectric", "metal", "plastic", "ggx", "beckmann", "checker"])
def test_furnace_per_material(pt, which):
    mats = {"lambert": pt.lambertian((1, 1, 1)), "dielectric": pt.dielectric(1.5), "metal": pt.metal((1, 1, 1), 0.4),
            "plastic": pt.plastic((1, 1, 1), 1.5), "ggx": pt.rough_conductor((1, 1, 1), 0.3, "al", pt.DIST_GGX),
            "beckmann": pt.rough_conductor((1, 1, 1), 0.25, "cu", pt.DIST_BECKMANN),
            "checker": pt.checker((1, 1, 1), (1, 1, 1), 0.5)}
    s = _furnace_scene(pt, mats[which])
    w = h = 64
    spp, depth = 64, 64
    cs = s.to_core().commit(0)
    g, _ = cs.render(s.camera, s.render_settings(width=w, height=h, spp=spp, max_depth=depth, seed=1))
    o, _ = OracleScene(s).render(s.camera, w, h, spp, depth, rng_mode=RNG_CHACHA)
    yy, xx = np.mgrid[0:h, 0:w]
    on_sphere = ((xx - w / 2 + 0.5) ** 2 + (yy - h / 2 + 0.5) ** 2) < (0.8 * w / 2 * np.tan(np.arcsin(1 / 4)) / np.tan(np.radians(15))) ** 2
    assert on_sphere.sum() > 300
    gm, om = g[on_sphere].mean(), o[on_sphere].mean()
    assert abs(gm - om) <= 0.005 * om, (which, gm, om)
    if which in ("lambert", "dielectric", "checker"):
        assert abs(gm - 0.5) < 1e-3 and g.min() > 0.499 and g.max() < 0.501  # KA5: energy conserving -> background
    else:
        # these can only lose energy (material.rs:106-108, tungsten/materials.rs:36,349-351); Metal loses it at grazing
        # angles only, which the 0.8-radius mask excludes
        assert gm <= 0.5 + 1e-6 and (which == "metal" or gm < 0.499)


def test_sharding_is_invisible(pt):
    s = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    cs = s.to_core().commit(0)
    w, h, spp, depth = 100, 70, 6, 6  # neither a multiple of 32
    base = dict(width=w, height=h, spp=spp, max_depth=depth, seed=4)
    whole, st = cs.render(s.camera, s.render_settings(**base))
    assert st.paths == w * h * spp
    by_sample = sum(cs.render(s.camera, s.render_settings(sample_begin=b, sample_end=e, **base))[0] for b, e in [(0, 1), (1, 4), (4, 6)])
    parts = [cs.render(s.camera, s.render_settings(tile_mod=3, tile_rem=r, **base)) for r in range(3)]
    by_tile = sum(p[0] for p in parts)
    assert sum(p[1].paths for p in parts) == w * h * spp
    assert np.allclose(by_sample, whole, rtol=1e-5, atol=1e-6)
    assert np.allclose(by_tile, whole, rtol=1e-5, atol=1e-6)
    # a tile shard touches only its own tiles
    tiles_x = (w + 31) // 32
    ys, xs = np.mgrid[0:h, 0:w]
    mine = ((ys // 32) * tiles_x + xs // 32) % 3 == 1
    assert (parts[1][0][~mine] == 0).all() and (parts[1][0][mine] > 0).any()
    # pool size and seed: the image must not depend on the pool, and must depend on the seed
    # the film is accumulated in fixed point: the image is independent of the order in which paths end, bit for bit
    small, _ = cs.render(s.camera, s.render_settings(pool_paths=2048, **base))
    assert np.array_equal(small, whole)
    again, _ = cs.render(s.camera, s.render_settings(**base))
    assert np.array_equal(again, whole)
    # a pool far larger than the job: every block takes a fair share of the paths, not the first blocks everything
    big, st_big = cs.render(s.camera, s.render_settings(pool_paths=1 << 21, **base))
    assert np.array_equal(big, whole) and st_big.rays == st.rays
    other, _ = cs.render(s.camera, s.render_settings(**dict(base, seed=5)))
    assert not np.allclose(other, whole, rtol=1e-3, atol=1e-4)


def test_film_nan_and_inf_survive_the_fixed_point_film(pt):
    # an emitter whose radiance exceeds the fixed-point range (and an infinite one) must come out as +inf, as an fp32 sum
    # would hold it — resolve then clamps it to 255 (renderer.rs:112-120) — not as a wrapped integer
    for value in (3e7, float("inf")):
        s = pt.Scene()
        m = s.add_material(pt.emissive((value, 1.0, 0.0)))
        s.add_sphere((0, 0, 0), 1.0, m)
        s.set_camera((0, 0, 4), (0, 0, 0), (0, 1, 0), 30.0, 1.0)
        cs = s.to_core().commit(0)
        img, _ = cs.render(s.camera, pt.RenderSettings(width=16, height=16, spp=4, max_depth=3))
        c = img[8, 8]
        assert np.isposinf(c[0]) and c[1] == 1.0 and c[2] == 0.0, c
        assert (cs.resolve_u32(img)[8 * 16 + 8] >> 16) == 255
        assert img[0, 0, 0] == 0.5  # background pixel untouched


def test_edge_cases(pt):
    # empty scene -> every path misses -> GRAY 0.5 -> 0xB4B4B4 (KA1)
    s = pt.Scene()
    s.set_camera((0, 0, 0), (0, 0, -1), (0, 1, 0), 60.0, 1.0)
    cs = s.to_core().commit(0)
    img, st = cs.render(s.camera, pt.RenderSettings(width=33, height=5, spp=3, max_depth=4))
    assert (img == 0.5).all() and st.rays == st.paths == 33 * 5 * 3
    assert (cs.resolve_u32(img) == 0x00B4B4B4).all()
    # max_depth 0: trace_ray returns black at once (renderer.rs:20-22); 1: camera segment only
    img0, st0 = cs.render(s.camera, pt.RenderSettings(width=8, height=8, spp=2, max_depth=0))
    assert (img0 == 0).all() and st0.rays == 0
    c = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    cc = c.to_core().commit(0)
    img1, st1 = cc.render(c.camera, c.render_settings(width=64, height=64, spp=1, max_depth=1))
    ref1, _ = OracleScene(c).render(c.camera, 64, 64, 1, 1, rng_mode=RNG_PHILOX)
    assert st1.rays == st1.paths == 64 * 64 and np.array_equal(img1, ref1)
    assert (img1.reshape(-1, 3).max(axis=1) > 0).mean() < 0.1  # only pixels that see the light directly


def test_hdr_sky_lookup(pt):
    rng = np.random.default_rng(0)
    sky = rng.uniform(0, 4, size=(16, 32, 3)).astype(np.float32)
    s = pt.Scene()
    m = s.add_material(pt.rough_conductor((1, 1, 1), 0.2, "ag", pt.DIST_GGX))
    s.add_sphere((0, 0, 0), 1.0, m)
    s.set_sky(sky)
    s.set_camera((0, 1, 4), (0, 0, 0), (0, 1, 0), 50.0, 1.5)
    cs = s.to_core().commit(0)
    g, gs = cs.render(s.camera, s.render_settings(width=96, height=64, spp=4, max_depth=6, seed=2))
    o, os_ = OracleScene(s).render(s.camera, 96, 64, 4, 6, rng_mode=RNG_PHILOX, seed=2)
    assert np.isclose(g, o, rtol=1e-3, atol=1e-4).all(axis=2).mean() > 0.99  # acos/atan2 differ by ulps at texel borders
    assert gs.rays == os_.rays


def test_device_resident_accumulate_and_resolve(pt):
    torch = pytest.importorskip("torch")
    s = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    cs = s.to_core().commit(0)
    w = h = 64
    st = s.render_settings(width=w, height=h, spp=4, max_depth=5, seed=8, flags=pt.FLAG_TIMING | pt.FLAG_COUNTERS)
    accum = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda:0")
    stats = cs.render_accumulate(s.camera, st, accum.data_ptr(), torch.cuda.current_stream().cuda_stream)
    stats2 = cs.render_accumulate(s.camera, st, accum.data_ptr(), torch.cuda.current_stream().cuda_stream)  # ADDS
    host, _ = cs.render(s.camera, st)
    assert np.allclose(accum.cpu().numpy().reshape(h, w, 3) / 8.0, host, rtol=1e-5, atol=1e-6)
    assert stats.extend_launches >= stats.iterations > 0 and stats.extend_ms > 0 and stats.shade_ms > 0
    assert stats.mesh_rays == 0 and stats2.rays == stats.rays
    out = torch.zeros(w * h, dtype=torch.int32, device="cuda:0")
    assert pt.core().ptc_resolve_device(accum.data_ptr(), w * h, 1.0 / 8.0, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert (out.cpu().numpy().view(np.uint32) == oracle_resolve(accum.cpu().numpy().reshape(-1, 3) / np.float32(8.0))).all()


def test_render_scene_entry_point_and_png(pt, tmp_path):
    # `render_scene(&scene, &camera, &settings) -> Vec<u32>` then `save_image` (main.rs:57-58)
    s = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    s.set_settings(48, 48, 4, 6)
    out = np.zeros(48 * 48, np.uint32)
    stats = pt.Stats()
    assert pt.host().pth_render_scene(s._h, 0, out.ctypes.data, C.byref(stats)) == 0
    assert stats.paths == 48 * 48 * 4 and (out >> 24 == 0).all() and out.max() > 0
    buf, img, _ = pt.render_scene(s, 0)
    assert np.array_equal(buf, out)  # seed 0 both times: the fixed-point film makes render_scene deterministic per seed
    pt.save_image(str(tmp_path / "c.png"), buf, 48, 48)
    assert os.path.getsize(tmp_path / "c.png") > 100


def test_render_u32_is_render_then_resolve(pt):
    s = pt.load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
    cs = s.to_core().commit(0)
    st = s.render_settings(width=96, height=64, spp=4, max_depth=6, seed=9)
    img, a = cs.render(s.camera, st)
    packed, b = cs.render_u32(s.camera, st)
    assert a.rays == b.rays and packed.shape == (96 * 64,)
    want = cs.resolve_u32(img)
    assert np.array_equal(packed, want)


def test_commit_twice_and_bad_device(pt):
    s = pt.Scene()
    m = s.add_material(pt.lambertian((1, 1, 1)))
    s.add_sphere((0, 0, 0), 1.0, m)
    cs = s.to_core()
    with pytest.raises(pt.PtcError) as e:
        cs.commit(99)
    assert e.value.code == pt.PTC_E_INVALID
    cs.commit(0)
    with pytest.raises(pt.PtcError) as e:
        cs.commit(0)
    assert e.value.code == pt.PTC_E_STATE
