#!/usr/bin/env python3
"""The command the ncu captures under profiles/ are taken on, and the quick timing probe used while tuning: config C2's
frame (semesterbild 800x600, depth 30) at a reduced sample count.  Same kernels, same launch shapes per iteration as
bench.py.

  python tools/profile_cmd.py [spp] [pool] [scene] [repeats] [flags]
     repeats = 1 : one render, no warm-up (what ncu wraps)
     repeats > 1 : warm-up + best-of-N with the per-stage CUDA-event split
     flags       : ptc_render_settings.flags; 2 (default) = per-stage kernels + event split, 0 = production path"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
name = sys.argv[3] if len(sys.argv) > 3 else "semesterbild.json"
repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 1
flags = int(sys.argv[5]) if len(sys.argv) > 5 else pt.FLAG_TIMING  # 0 = production path (one persistent kernel)
if name == "synthetic":
    scene = pt.synthetic_scene(cells=int(os.environ.get("CELLS", "1000")))
    scene.set_settings(1920, 1080, spp, 16)
else:
    scene = pt.load_scene_from_json(os.path.join(ROOT, "scenes", name))
cs = scene.to_core().commit(0)
st = scene.render_settings(spp=spp, seed=0, pool_paths=pool, flags=flags)
if repeats > 1:
    cs.render(scene.camera, st)
best = None
for _ in range(repeats):
    img, stats = cs.render(scene.camera, st)
    if best is None or stats.render_ms < best.render_ms:
        best = stats
s = best
other = s.render_ms - s.extend_ms - s.shade_ms - s.regen_ms
print(f"{name} {spp}spp pool={pool}: {s.render_ms:.2f} ms  pre {s.pre_ms:.2f}  traverse {s.traverse_ms:.2f}  post {s.post_ms:.2f}  "
      f"shade {s.shade_ms:.2f}  regen {s.regen_ms:.2f}  other {other:.2f} | {s.paths / s.render_ms / 1e3:.0f} Mpaths/s "
      f"{s.rays / s.render_ms / 1e3:.0f} Mrays/s  iters {s.iterations} launches {s.kernel_launches} rays {s.rays}")
