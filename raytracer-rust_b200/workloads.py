"""The five BASELINE.json configurations (SURVEY.md 8d "Concrete inputs") as host scene descriptions at their NATIVE
resolution / spp / bounce settings — one definition shared by bench.py (`--config`), the parity tests and the tools.

  C1  Cornell box           scenes/cornell-box/scene.json with the overrides 256x256, 16 spp, depth 8
  C2  semesterbild.json     as shipped: 800x600, 256 spp, depth 30 (3 cubes, 4,748-triangle mesh, glass sphere; sky-lit)
  C3  teapot (derived)      the shipped Tungsten teapot JSON cannot be loaded by the reference (`infinite_sphere` is not an
                            ObjectConfigVariant, src/tungsten/parser.rs:136-165; WO3 mis-read, mesh_object.rs:188-190), so:
                            teapot.obj x30, rough_conductor GGX copper, checker floor quad; 1280x720, 512 spp, depth 64
  C3s the shipped teapot    loaded with the loader extensions (sky from infinite_sphere, 16-byte WO3 records)
  C4  veach-mis             as shipped: 1280x720, 1024 spp, depth 16 (Beckmann rough conductors, 3 emissive spheres)
  C5  synthetic             1000x1000-cell height field = 2,000,000 triangles + glass sphere + GGX-Al cube + emissive quad;
                            3840x2160, 256 spp, depth 16
"""
import os

from . import (DIST_GGX, LOAD_INFINITE_SPHERE_SKY, LOAD_WO3_STRIDE16, REPO_ROOT, Scene, checker, load_scene_from_json,
               rough_conductor, synthetic_scene)

SCENES = os.path.join(REPO_ROOT, "scenes")
NAMES = ("C1", "C2", "C3", "C3s", "C4", "C5")


def teapot_scene():
    s = Scene()
    cu = s.add_material(rough_conductor((1, 1, 1), 0.1, "cu", DIST_GGX))
    ck = s.add_material(checker((0.8, 0.8, 0.8), (0.2, 0.2, 0.2), 40.0))
    s.add_obj(os.path.join(SCENES, "teapot", "teapot.obj"), cu, scale=(30, 30, 30), rotation=(0, 30, 0), position=(0, 0, 0))
    s.add_quad(ck, scale=(900, 1, 900), rotation=(0, 0, 180), position=(0, 0, 0))
    s.set_camera((0, 130, 330), (0, 45, 0), (0, 1, 0), 35.0, 1280 / 720)
    s.set_settings(1280, 720, 512, 64)
    return s


def workload(name, cells=1000):
    """-> (description, Scene) with the scene's settings at the configuration's native values."""
    if name == "C1":
        s = load_scene_from_json(os.path.join(SCENES, "cornell-box", "scene.json"))
        s.set_settings(256, 256, 16, 8)
        return "C1 cornell-box/scene.json 256x256, 16 spp, max_bounces 8 (6 quads + 2 cubes, lambert; one emissive quad)", s
    if name == "C2":
        s = load_scene_from_json(os.path.join(SCENES, "semesterbild.json"))
        return "C2 semesterbild.json 800x600, 256 spp, max_bounces 30 (3 cubes, 4748-triangle mesh, glass sphere; sky-lit)", s
    if name == "C3":
        return "C3 teapot (derived: teapot.obj x30, GGX copper, checker quad) 1280x720, 512 spp, max_bounces 64", teapot_scene()
    if name == "C3s":
        s = load_scene_from_json(os.path.join(SCENES, "teapot", "scene.json"), LOAD_INFINITE_SPHERE_SKY | LOAD_WO3_STRIDE16)
        s.set_settings(1280, 720, 512, 64)
        return "C3s tungsten/teapot/scene.json as shipped (loader extensions: HDR sky, WO3 stride 16) 1280x720, 512 spp, max_bounces 64", s
    if name == "C4":
        s = load_scene_from_json(os.path.join(SCENES, "veach-mis", "scene.json"))
        return "C4 veach-mis/scene.json 1280x720, 1024 spp, max_bounces 16 (4 Beckmann-Cu cubes, 2 quads, 3 emissive spheres)", s
    if name == "C5":
        s = synthetic_scene(cells=cells)
        n = 2 * cells * cells
        return f"C5 synthetic {n:,}-triangle height field + glass sphere + GGX-Al cube + emissive quad 3840x2160, 256 spp, max_bounces 16", s
    raise ValueError(f"unknown config {name!r} (one of {NAMES})")
