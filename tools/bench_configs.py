#!/usr/bin/env python3
"""All five BASELINE.json configurations on one B200, next to the CPU oracle on the same box: Mpaths/s, Mrays/s, the
per-stage split (CUDA events), traversal counters and the extend roofline.  bench.py stays the headline
(C2 at full size, the contract's JSON line); this tool is the per-config table behind DESIGN.md / profiles/.

  python tools/bench_configs.py [--out gpurun_out/configs.json] [--cpu-seconds 4]

Sample counts are reduced where the full count would only repeat the same work (throughput does not depend on spp once
the path pool is saturated); resolutions and bounce limits are the configs' own.
"""
import argparse
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
from bindings import RNG_CHACHA, OracleScene  # noqa: E402

SC = os.path.join(ROOT, "scenes")


def teapot_scene():
    """C3 (derived): the shipped Tungsten teapot JSON cannot be loaded by the reference (`infinite_sphere`
    primitive, mis-read WO3 meshes), so: data/models/teapot.obj scaled x30 (its tiny triangles otherwise fall under the
    |a| < 1e-4 cull), rough_conductor GGX copper, checker floor quad, default grey background."""
    s = pt.Scene()
    cu = s.add_material(pt.rough_conductor((1, 1, 1), 0.1, "cu", pt.DIST_GGX))
    ck = s.add_material(pt.checker((0.8, 0.8, 0.8), (0.2, 0.2, 0.2), 40.0))
    s.add_obj(os.path.join(SC, "teapot", "teapot.obj"), cu, scale=(30, 30, 30), rotation=(0, 30, 0), position=(0, 0, 0))
    s.add_quad(ck, scale=(900, 1, 900), rotation=(0, 0, 180), position=(0, 0, 0))
    s.set_camera((0, 130, 330), (0, 45, 0), (0, 1, 0), 35.0, 1280 / 720)
    s.set_settings(1280, 720, 512, 64)
    return s


def configs():
    c1 = pt.load_scene_from_json(os.path.join(SC, "cornell-box", "scene.json"))
    c1.set_settings(256, 256, 16, 8)
    yield "C1 cornell-box 256x256x16spp depth 8", c1, 16, 16
    c2 = pt.load_scene_from_json(os.path.join(SC, "semesterbild.json"))
    yield "C2 semesterbild 800x600x256spp depth 30", c2, 256, 8
    yield "C3 teapot (derived) 1280x720x512spp depth 64", teapot_scene(), 64, 2
    c3s = pt.load_scene_from_json(os.path.join(SC, "teapot", "scene.json"), pt.LOAD_INFINITE_SPHERE_SKY | pt.LOAD_WO3_STRIDE16)
    c3s.set_settings(1280, 720, 512, 64)
    yield "C3s teapot as shipped (loader extensions: sky, WO3) 1280x720x512spp depth 64", c3s, 64, 2
    c4 = pt.load_scene_from_json(os.path.join(SC, "veach-mis", "scene.json"))
    yield "C4 veach-mis 1280x720x1024spp depth 16", c4, 128, 4
    c5 = pt.synthetic_scene(cells=1000)
    yield "C5 synthetic 2M-triangle height field 3840x2160x256spp depth 16", c5, 16, 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    peak_hbm = 6540.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak_hbm = float(json.load(open(p))["hbm_gbs"])
    cost = {pt.OBJ_SPHERE: 25, pt.OBJ_PLANE: 14, pt.OBJ_QUAD: 35, pt.OBJ_CUBE: 80, pt.OBJ_MESH: 64}
    rows = []
    for name, scene, gpu_spp, cpu_spp in configs():
        w, h, spp, depth = scene.settings
        t0 = time.time()
        cs = scene.to_core().commit(0)
        commit_s = time.time() - t0
        cam = scene.camera
        st = scene.render_settings(spp=gpu_spp, seed=0, pool_paths=1 << 23)
        cs.render(cam, st)  # warm-up
        best = None
        for _ in range(3):
            _, s = cs.render(cam, st)
            if best is None or s.render_ms < best.render_ms:
                best = s
        st.flags = pt.FLAG_TIMING
        _, tm = cs.render(cam, st)
        st.flags = pt.FLAG_COUNTERS
        st.spp = min(gpu_spp, 8)
        _, ct = cs.render(cam, st)
        npr, tpr = ct.nodes_visited / ct.rays, ct.tris_tested / ct.rays
        bytes_per_ray = 48 + 80.0 * npr + 48.0 * tpr
        instr_per_ray = sum(cost[o.type] for o in scene.objects) + 170.0 * npr + 45.0 * tpr
        ext_s = tm.extend_ms / 1e3
        row = {
            "config": name, "gpu_spp_run": gpu_spp, "width": w, "height": h, "depth": depth, "commit_s": commit_s,
            "mpaths_s": best.paths / best.render_ms / 1e3, "mrays_s": best.rays / best.render_ms / 1e3,
            "rays_per_path": best.rays / best.paths, "render_ms": best.render_ms, "iterations": best.iterations,
            "stage_ms": {"pre": tm.pre_ms, "traverse": tm.traverse_ms, "post": tm.post_ms, "shade": tm.shade_ms, "total": tm.render_ms},
            "nodes_per_ray": npr, "tris_per_ray": tpr, "algorithmic_bytes_per_ray": bytes_per_ray,
            "algorithmic_instr_per_ray": instr_per_ray,
            "extend_gbs": bytes_per_ray * tm.rays / ext_s / 1e9, "extend_frac_of_hbm": bytes_per_ray * tm.rays / ext_s / 1e9 / peak_hbm,
            "extend_grays_s": tm.rays / ext_s / 1e9,
        }
        if not args.no_cpu:
            orc = OracleScene(scene)
            _, os_ = orc.render(cam, w, h, cpu_spp, depth, rng_mode=RNG_CHACHA)
            row["cpu"] = {"mpaths_s": os_.paths / os_.seconds / 1e6, "mrays_s": os_.rays / os_.seconds / 1e6, "spp_run": cpu_spp,
                          "seconds": os_.seconds, "cores": os.cpu_count()}
            row["speedup_paths"] = row["mpaths_s"] / row["cpu"]["mpaths_s"]
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    print("\n%-62s %10s %10s %6s %8s %8s %9s" % ("config", "Mpaths/s", "Mrays/s", "r/p", "CPU Mp/s", "speedup", "ext %HBM"))
    for r in rows:
        print("%-62s %10.0f %10.0f %6.2f %8.2f %8.0f %9.1f" % (r["config"][:62], r["mpaths_s"], r["mrays_s"], r["rays_per_path"],
                                                             r.get("cpu", {}).get("mpaths_s", float("nan")), r.get("speedup_paths", float("nan")),
                                                             100 * r["extend_frac_of_hbm"]))


if __name__ == "__main__":
    main()
