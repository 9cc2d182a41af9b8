#!/usr/bin/env python3
"""How much faster does the traversal run on COHERENT rays?  Config C5's scene traced with max_depth = 1 (camera rays only:
32 adjacent pixels per warp) against its usual depth 16 (rays in shade order): traversal time per node step.  An upper
bound for what sorting the parked rays spatially could buy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
scene = pt.synthetic_scene(cells=int(os.environ.get("CELLS", "1000")))
scene.set_settings(1920, 1080, 16, 16)
cs = scene.to_core().commit(0)
for depth in (1, 2, 16):
    st = scene.render_settings(spp=16, max_depth=depth, seed=0, pool_paths=1 << 23, flags=pt.FLAG_TIMING)
    cs.render(scene.camera, st)
    _, s = cs.render(scene.camera, st)
    _, c = cs.render(scene.camera, scene.render_settings(spp=4, max_depth=depth, seed=0, pool_paths=1 << 23, flags=pt.FLAG_COUNTERS))
    nodes_per_ray = c.nodes_visited / max(1, c.rays)
    steps = nodes_per_ray * s.rays
    print(f"depth {depth:2d}: rays {s.rays:10d}  traverse {s.traverse_ms:7.3f} ms  nodes/ray {nodes_per_ray:5.2f}  "
          f"-> {s.traverse_ms * 1e6 / steps:6.3f} ns per node step  (pre {s.pre_ms:.2f} post {s.post_ms:.2f} shade {s.shade_ms:.2f})")
