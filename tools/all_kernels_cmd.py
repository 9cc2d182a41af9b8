#!/usr/bin/env python3
"""A small workload that touches every kernel (written for compute-sanitizer, which turned out to be closed on this
GPU pool; still the quickest whole-library run):
semesterbild shrunk (mesh: pre / traverse with stealing / post / shade), the Cornell box with NEE (k_shade<true>, shadow
rays), ptc_intersect, and the device-side mesh build."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
S = os.path.join(ROOT, "scenes")
s = pt.load_scene_from_json(os.path.join(S, "semesterbild.json"))
cs = s.to_core().commit(0)
img, st = cs.render(s.camera, s.render_settings(width=96, height=72, spp=2, max_depth=12, seed=1, pool_paths=8192))
print("semesterbild", st.paths, st.rays, float(img.mean()))
o, d = cs.primary_rays(s.camera, s.render_settings(width=64, height=48), 0)
h, _ = cs.intersect(o, d)
print("intersect hits", int((h["object"] >= 0).sum()))
c = pt.load_scene_from_json(os.path.join(S, "cornell-box", "scene.json"))
cc = c.to_core().commit(0)
img, st = cc.render(c.camera, c.render_settings(width=64, height=64, spp=2, max_depth=6, seed=1, flags=pt.FLAG_NEE, pool_paths=4096))
print("cornell nee", st.paths, st.rays, float(img.mean()))
os.environ["PTC_BUILD"] = "device"
t = pt.synthetic_scene(cells=48)
ct = t.to_core().commit(0)
img, st = ct.render(t.camera, t.render_settings(width=64, height=36, spp=1, max_depth=4, seed=1))
print("device-built synthetic", st.paths, st.rays, float(img.mean()))
