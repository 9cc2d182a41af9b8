"""Next-event estimation + multiple importance sampling (PTC_FLAG_NEE) — an integrator option the
reference does not have (its IntegratorConfig ignores "enable_mis", src/tungsten/parser.rs:167-171).  It must not change
WHAT is computed, only how fast it converges: the expected image is the plain integrator's, whose parity with the
reference the other tests establish.

  * unbiasedness: image means (global and per 8x8 block) of an NEE render agree with a plain render of much higher
    sample count within the Monte-Carlo error of the two;
  * variance: at equal spp the NEE render is closer (relMSE) to that high-spp plain render than the plain render is —
    by a large factor on veach-mis (small emissive spheres) and the Cornell box (one quad light);
  * off by default: flags = 0 renders are bit-identical to a build that never heard of NEE (the same-stream oracle tests);
  * deterministic: same seed -> same image, bit for bit; scenes without sampled lights are untouched by the flag.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "scenes")


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def blocks(img, k=8):
    h, w, _ = img.shape
    return img[:h // k * k, :w // k * k].reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))


@pytest.mark.parametrize("name,w,h,depth,ref_spp,spp", [("cornell-box/scene.json", 96, 96, 8, 4096, 64),
                                                         ("veach-mis/scene.json", 160, 90, 16, 8192, 64)])
def test_nee_is_unbiased_and_converges_faster(pt, name, w, h, depth, ref_spp, spp):
    s = pt.load_scene_from_json(os.path.join(SCENES, name))
    cs = s.to_core().commit(0)
    base = dict(width=w, height=h, max_depth=depth)
    ref, _ = cs.render(s.camera, s.render_settings(spp=ref_spp, seed=1, **base))                      # plain, converged
    plain, ps = cs.render(s.camera, s.render_settings(spp=spp, seed=2, **base))
    nee, ns = cs.render(s.camera, s.render_settings(spp=spp, seed=2, flags=pt.FLAG_NEE, **base))
    nee_hi, _ = cs.render(s.camera, s.render_settings(spp=spp * 16, seed=3, flags=pt.FLAG_NEE, **base))
    assert ns.paths == ps.paths and ns.rays > ps.rays  # shadow rays are extend items too
    # unbiased: per-channel global means, and 8x8 block means of the higher-spp NEE render, against the converged plain render
    for c in range(3):
        assert abs(nee_hi[..., c].mean() - ref[..., c].mean()) <= 0.01 * ref[..., c].mean(), (c, nee_hi[..., c].mean(), ref[..., c].mean())
    bn, br = blocks(nee_hi), blocks(ref)
    rel = np.abs(bn - br) / (br + 1e-2)
    assert np.median(rel) < 0.02 and np.percentile(rel, 99) < 0.15, (np.median(rel), np.percentile(rel, 99))
    # variance: at equal spp NEE is much closer to the truth
    e_plain, e_nee = relmse(plain, ref), relmse(nee, ref)
    assert e_nee < 0.5 * e_plain, (e_nee, e_plain)
    # deterministic
    again, _ = cs.render(s.camera, s.render_settings(spp=spp, seed=2, flags=pt.FLAG_NEE, **base))
    assert np.array_equal(again, nee)


def test_nee_without_sampled_lights_changes_nothing(pt):
    # semesterbild has no emitter at all (lit by the background): the flag must leave the image bit-identical
    s = pt.load_scene_from_json(os.path.join(SCENES, "semesterbild.json"))
    cs = s.to_core().commit(0)
    base = dict(width=160, height=120, spp=4, max_depth=30, seed=5)
    a, sa = cs.render(s.camera, s.render_settings(**base))
    b, sb = cs.render(s.camera, s.render_settings(flags=pt.FLAG_NEE, **base))
    assert np.array_equal(a, b) and sa.rays == sb.rays


def test_nee_all_continuous_materials(pt):
    # plastic, GGX and Beckmann rough conductors, checker and plain Lambertian under one small sphere light + a quad
    # light: block means of the NEE render against a converged plain render
    s = pt.Scene()
    floor = s.add_material(pt.checker((0.8, 0.8, 0.8), (0.3, 0.3, 0.3), 2.0))
    mats = [pt.lambertian((0.7, 0.3, 0.3)), pt.plastic((0.2, 0.6, 0.3), 1.5), pt.rough_conductor((1, 1, 1), 0.3, "cu", pt.DIST_GGX),
            pt.rough_conductor((1, 1, 1), 0.25, "al", pt.DIST_BECKMANN), pt.dielectric(1.5), pt.metal((0.9, 0.9, 0.9), 0.2)]
    s.add_quad(floor, scale=(40, 1, 40), rotation=(0, 0, 180), position=(0, 0, 0))
    for k, m in enumerate(mats):
        s.add_sphere((-7.5 + 3.0 * k, 1.0, 0.0), 1.0, s.add_material(m))
    s.add_sphere((0.0, 6.0, 2.0), 0.4, s.add_material(pt.emissive((60, 50, 40))))
    s.add_quad(s.add_material(pt.emissive((4, 5, 6))), scale=(3, 1, 3), rotation=(0, 0, 0), position=(-4, 8, -2))
    s.set_camera((0, 5, 14), (0, 1, 0), (0, 1, 0), 40.0, 2.0)
    cs = s.to_core().commit(0)
    base = dict(width=128, height=64, max_depth=6)
    ref, _ = cs.render(s.camera, s.render_settings(spp=16384, seed=1, **base))
    nee, _ = cs.render(s.camera, s.render_settings(spp=2048, seed=2, flags=pt.FLAG_NEE, **base))
    plain, _ = cs.render(s.camera, s.render_settings(spp=2048, seed=2, **base))
    bn, br = blocks(nee), blocks(ref)
    rel = np.abs(bn - br) / (br + 1e-2)
    assert np.median(rel) < 0.02 and np.percentile(rel, 99) < 0.12, (np.median(rel), np.percentile(rel, 99))
    for c in range(3):
        assert abs(nee[..., c].mean() - ref[..., c].mean()) <= 0.01 * ref[..., c].mean()
    assert relmse(nee, ref) < relmse(plain, ref)
