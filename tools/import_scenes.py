#!/usr/bin/env python3
"""Import the reference's scene *data* (inputs, not source code) into scenes/.

The GPU box has no /root/reference, so the scene descriptions the benchmark is
quoted on (BASELINE.json configs) are carried as fixtures.  JSON files are
re-serialised (semantically identical, key order kept); OBJ meshes are copied
byte for byte because the loader's parse of them is part of what is tested.

  python tools/import_scenes.py [/root/reference]

Provenance (all relative to the reference root):
  data/scenes/semesterbild.json                 -> scenes/semesterbild.json        (config C2)
  data/scenes/RayTracingText.obj                -> scenes/RayTracingText.obj
  data/scenes/tungsten/cornell-box/scene.json   -> scenes/cornell-box/scene.json   (config C1)
  data/scenes/tungsten/veach-mis/scene.json     -> scenes/veach-mis/scene.json     (config C4)
  data/models/teapot.obj                        -> scenes/teapot/teapot.obj        (config C3, derived scene)
  data/scenes/tungsten/teapot/{scene.json, models/*.wo3, textures/envmap.hdr} -> scenes/teapot/   (the shipped teapot
                                                   scene: fails to load in the reference; loads with the LOAD_* extensions)
"""
import json
import os
import shutil
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")

JSONS = [
    ("data/scenes/semesterbild.json", "semesterbild.json"),
    ("data/scenes/tungsten/cornell-box/scene.json", "cornell-box/scene.json"),
    ("data/scenes/tungsten/veach-mis/scene.json", "veach-mis/scene.json"),
    ("data/scenes/tungsten/teapot/scene.json", "teapot/scene.json"),
]
COPIES = [
    ("data/scenes/RayTracingText.obj", "RayTracingText.obj"),
    ("data/models/teapot.obj", "teapot/teapot.obj"),
    ("data/scenes/tungsten/teapot/models/Mesh000.wo3", "teapot/models/Mesh000.wo3"),
    ("data/scenes/tungsten/teapot/models/Mesh001.wo3", "teapot/models/Mesh001.wo3"),
    ("data/scenes/tungsten/teapot/textures/envmap.hdr", "teapot/textures/envmap.hdr"),
]

for src, dst in JSONS:
    with open(os.path.join(ref, src)) as f:
        doc = json.load(f)
    out = os.path.join(root, dst)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print("json ", src, "->", dst)

for src, dst in COPIES:
    out = os.path.join(root, dst)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    shutil.copyfile(os.path.join(ref, src), out)
    print("copy ", src, "->", dst)
