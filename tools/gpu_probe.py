#!/usr/bin/env python3
"""Exploratory run on a B200 box: parity against the oracle + first timings.  Writes gpurun_out/probe.log.

  python tools/gpu_probe.py [quick]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ptload  # noqa: E402

pt = ptload.load()
from bindings import OracleScene, compare_hits, oracle_scatter  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")


def say(**kw):
    line = json.dumps(kw, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    say(what="devices", n=pt.device_count())
    try:
        import torch
        p = torch.cuda.get_device_properties(0)
        say(what="device", name=p.name, sms=p.multi_processor_count, l2=getattr(p, "L2_cache_size", None), mem=p.total_memory)
    except Exception as e:  # noqa: BLE001
        say(what="device", error=str(e))

    # ---- closest-hit parity on primary rays
    for name, over in [("semesterbild.json", {}), ("cornell-box/scene.json", dict(width=256, height=256)),
                       ("veach-mis/scene.json", dict(width=640, height=360))]:
        s = pt.load_scene_from_json(os.path.join(ROOT, "scenes", name))
        st = s.render_settings(**over)
        cs = s.to_core().commit(0)
        orc = OracleScene(s)
        o, d = cs.primary_rays(s.camera, st, 0)
        got, stats = cs.intersect(o, d)
        want = orc.intersect(pt, o, d)
        r = compare_hits(got, want)
        r.pop("id_bad_idx")
        say(what="primary_parity", scene=name, **r, ms=stats.render_ms, nodes_per_mesh_ray=stats.nodes_visited / max(1, stats.mesh_rays),
            tris_per_mesh_ray=stats.tris_tested / max(1, stats.mesh_rays))

    # ---- scatter parity per material
    rng = np.random.default_rng(3)
    n = 20000
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    flip = (dirs * nrm).sum(1) > 0
    dirs[flip] *= -1  # hit normals oppose the ray (set_face_normal)
    pos = rng.uniform(-50, 50, size=(n, 3))
    ff = rng.integers(0, 2, size=n)
    u4 = rng.uniform(0, 1, size=(n, 4)).astype(np.float32)
    mats = [pt.lambertian((0.7, 0.6, 0.5)), pt.checker((0.9, 0.9, 0.9), (0.1, 0.2, 0.3), 10.0), pt.metal((0.8, 0.8, 0.9), 0.3),
            pt.dielectric(1.52), pt.emissive((3, 2, 1)), pt.plastic((0.2, 0.5, 0.9), 1.5),
            pt.rough_conductor((0.2, 0.3, 0.6), 0.1, "al", pt.DIST_GGX), pt.rough_conductor((1, 1, 1), 0.05, "cu", pt.DIST_BECKMANN)]
    s = pt.Scene()
    for m in mats:
        s.add_material(m)
    s.add_sphere((0, 0, 0), 1.0, 0)
    cs = s.to_core().commit(0)
    for i, m in enumerate(mats):
        g = cs.scatter(i, dirs, pos, nrm, ff, u4)
        w = oracle_scatter(s.materials[i], dirs, pos, nrm, ff, u4)
        same = g[0] == w[0]
        both = same & (w[0] == 1)
        err_d = float(np.abs(g[2][both] - w[2][both]).max()) if both.any() else 0.0
        err_o = float(np.abs(g[1][both] - w[1][both]).max()) if both.any() else 0.0
        err_a = float((np.abs(g[3][both] - w[3][both]) / (np.abs(w[3][both]) + 1e-6)).max()) if both.any() else 0.0
        say(what="scatter_parity", material=i, type=m.type, flag_mismatch=int((~same).sum()), scattered=int(w[0].sum()),
            max_abs_dir=err_d, max_abs_origin=err_o, max_rel_att=err_a, emitted_eq=bool((g[4] == w[4]).all()))

    # ---- image parity, small
    s = pt.load_scene_from_json(os.path.join(ROOT, "scenes", "cornell-box", "scene.json"))
    cs = s.to_core().commit(0)
    orc = OracleScene(s)
    st = s.render_settings(width=128, height=128, spp=16, max_depth=8, seed=1)
    img, stats = cs.render(s.camera, st)
    ref, ostats = orc.render(s.camera, 128, 128, 16, 8, seed=1)
    close = np.isclose(img, ref, rtol=1e-3, atol=1e-4).all(axis=2).mean()
    say(what="image_parity", scene="cornell 128x128x16x8", gpu_mean=float(img.mean()), oracle_mean=float(ref.mean()),
        relmse=relmse(img, ref), pixels_close=float(close), gpu_rays=stats.rays, oracle_rays=ostats.rays, stats=stats.as_dict())

    s = pt.load_scene_from_json(os.path.join(ROOT, "scenes", "semesterbild.json"))
    cs = s.to_core().commit(0)
    orc = OracleScene(s)
    st = s.render_settings(width=200, height=150, spp=8, max_depth=30, seed=2)
    img, stats = cs.render(s.camera, st)
    t0 = time.time()
    ref, ostats = orc.render(s.camera, 200, 150, 8, 30, seed=2)
    say(what="image_parity", scene="semesterbild 200x150x8x30", gpu_mean=float(img.mean()), oracle_mean=float(ref.mean()),
        relmse=relmse(img, ref), pixels_close=float(np.isclose(img, ref, rtol=1e-3, atol=1e-4).all(axis=2).mean()),
        gpu_rays=stats.rays, oracle_rays=ostats.rays, oracle_s=time.time() - t0, oracle_mpaths=ostats.paths / ostats.seconds / 1e6,
        stats=stats.as_dict())

    # ---- timings on the headline config (reduced spp; throughput is spp-independent)
    spp = 16 if quick else 64
    for pool in ([1 << 20] if quick else [1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22]):
        st = s.render_settings(spp=spp, seed=0, pool_paths=pool, flags=pt.FLAG_TIMING)
        cs.render(s.camera, st)
        best = None
        for _ in range(3):
            img, stats = cs.render(s.camera, st)
            if best is None or stats.render_ms < best.render_ms:
                best = stats
        say(what="timing", scene="semesterbild 800x600", spp=spp, pool=pool, mpaths_s=best.paths / best.render_ms / 1e3,
            mrays_s=best.rays / best.render_ms / 1e3, rays_per_path=best.rays / best.paths, stats=best.as_dict())
    st = s.render_settings(spp=8, seed=0, flags=pt.FLAG_COUNTERS)
    img, stats = cs.render(s.camera, st)
    say(what="counters", scene="semesterbild 800x600x8", nodes_per_ray=stats.nodes_visited / stats.rays, tris_per_ray=stats.tris_tested / stats.rays,
        mesh_rays_per_ray=stats.mesh_rays / stats.rays, stats=stats.as_dict())
    u32 = cs.resolve_u32(img)
    pt.save_image(os.path.join(ROOT, "gpurun_out", "semesterbild_8spp.png"), u32, 800, 600)
    say(what="top_left_pixel", value=hex(int(u32[0])))


if __name__ == "__main__":
    main()
