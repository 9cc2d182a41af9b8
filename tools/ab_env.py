#!/usr/bin/env python3
"""A/B of a run-time switch of the core (an environment variable the library reads per call, e.g. PTC_STEAL, PTC_REFILL)
on one B200: a BASELINE config at its native settings and at the 1/8 sample share one of eight GPUs gets under strong
scaling.  One JSON line per (config, share, value): best of 6 after warm-up, CUDA-event render time.

  python tools/ab_env.py PTC_STEAL 1,0 C2 C5        (AB_POOL = pool slots, default 12 Mi; 0 = the library's default)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
from raytracer_rust_b200 import workloads  # noqa: E402


def main():
    var, values = sys.argv[1], sys.argv[2].split(",")
    for cfg in sys.argv[3:] or ["C2"]:
        label, scene = workloads.workload(cfg)
        cs = scene.to_core().commit(0)
        w, h, spp, depth = scene.settings
        for share in (1, 8):
            for v in values:
                os.environ[var] = v
                st = scene.render_settings(spp=spp, sample_begin=0, sample_end=max(1, spp // share), seed=0, pool_paths=int(os.environ.get("AB_POOL", str(3 << 22))))
                best = None
                for _ in range(6):
                    _, s = cs.render_u32(scene.camera, st)
                    if best is None or s.render_ms < best.render_ms:
                        best = s
                st.flags = pt.FLAG_TIMING
                _, tm = cs.render_u32(scene.camera, st)
                print(json.dumps({"config": cfg, "share": f"1/{share}", var: v, "render_ms": best.render_ms, "iterations": best.iterations,
                                  "launches": best.kernel_launches, "rays": best.rays, "mpaths_s": best.paths / best.render_ms / 1e3,
                                  "stage_ms": {"pre": tm.pre_ms, "traverse": tm.traverse_ms, "post": tm.post_ms, "shade": tm.shade_ms}}), flush=True)


if __name__ == "__main__":
    main()
