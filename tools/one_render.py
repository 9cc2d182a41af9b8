#!/usr/bin/env python3
"""One render of a BASELINE config (optionally a 1/N sample share of it) through ptc_render_u32 — the command ncu is pointed
at when single launches of the wavefront are to be captured (e.g. the near-empty launches of the drain:
`--launch-skip 100 --launch-count 32`).

  python tools/one_render.py C2 [share] [warmups]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
from raytracer_rust_b200 import workloads  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
share = int(sys.argv[2]) if len(sys.argv) > 2 else 1
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
label, scene = workloads.workload(cfg)
cs = scene.to_core().commit(0)
w, h, spp, depth = scene.settings
st = scene.render_settings(spp=spp, sample_begin=0, sample_end=max(1, spp // share), seed=0, pool_paths=int(os.environ.get("AB_POOL", str(3 << 22))))
for _ in range(warm + 1):
    _, s = cs.render_u32(scene.camera, st)
print(json.dumps({"config": label, "share": share, "render_ms": s.render_ms, "iterations": s.iterations, "launches": s.kernel_launches,
                  "rays": s.rays, "paths": s.paths}))
