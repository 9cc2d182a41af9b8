#!/usr/bin/env python3
"""The command the ncu captures under profiles/ are taken on: config C2's frame (semesterbild 800x600, depth 30) at a
reduced sample count so the capture stays short.  Same kernels, same launch shapes per iteration as bench.py.

  python tools/profile_cmd.py [spp] [pool] [scene]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
name = sys.argv[3] if len(sys.argv) > 3 else "semesterbild.json"
if name == "synthetic":
    scene = pt.synthetic_scene(cells=int(os.environ.get("CELLS", "1000")))
    scene.set_settings(1920, 1080, spp, 16)
else:
    scene = pt.load_scene_from_json(os.path.join(ROOT, "scenes", name))
cs = scene.to_core().commit(0)
st = scene.render_settings(spp=spp, seed=0, pool_paths=pool, flags=pt.FLAG_TIMING)
img, stats = cs.render(scene.camera, st)
print({k: v for k, v in stats.as_dict().items()}, "Mpaths/s", stats.paths / stats.render_ms / 1e3, "Mrays/s", stats.rays / stats.render_ms / 1e3)
