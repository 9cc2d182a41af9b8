#!/usr/bin/env python3
"""C2 at its native settings and at the 1/8 sample share for a range of pool sizes (16 … 160 Mi slots): where the default
pool policy of the library (whole job in flight, or 48 Mi, never 1–2.5 fills) comes from.
One JSON line per (share, pool)."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import ptload
pt = ptload.load()
from raytracer_rust_b200 import workloads
label, scene = workloads.workload("C2")
cs = scene.to_core().commit(0)
w, h, spp, depth = scene.settings
for share in (1, 8):
    for pool_mi in (16, 48, 96, 120, 128, 160):
        st = scene.render_settings(spp=spp, sample_begin=0, sample_end=spp // share, seed=0, pool_paths=pool_mi << 20)
        best = None
        for _ in range(5):
            _, s = cs.render_u32(scene.camera, st)
            if best is None or s.render_ms < best.render_ms:
                best = s
        print(json.dumps({"share": f"1/{share}", "pool_Mi": pool_mi, "render_ms": round(best.render_ms, 3), "iterations": best.iterations}), flush=True)
