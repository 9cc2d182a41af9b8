#!/usr/bin/env python3
"""A/B of the drain's tail threshold (PTC_TAIL_RAYS) on one B200: C2 at its native 256 spp and at the 32-spp share one of
eight GPUs gets under strong scaling.  Prints one JSON line per setting (best of 5 after warm-up, CUDA-event render time)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
from raytracer_rust_b200 import workloads  # noqa: E402


def main():
    cfgs = sys.argv[1:] or ["C2"]
    for cfg in cfgs:
        label, scene = workloads.workload(cfg)
        cs = scene.to_core().commit(0)
        w, h, spp, depth = scene.settings
        for share in (1, 8):
            for tail in (0, 4096, 16384, 65536, 262144):
                os.environ["PTC_TAIL_RAYS"] = str(tail)
                st = scene.render_settings(spp=spp, sample_begin=0, sample_end=max(1, spp // share), seed=0, pool_paths=3 << 22)
                best = None
                for _ in range(6):
                    _, s = cs.render_u32(scene.camera, st)
                    if best is None or s.render_ms < best.render_ms:
                        best = s
                print(json.dumps({"config": cfg, "share": f"1/{share}", "tail_rays": tail, "render_ms": best.render_ms, "iterations": best.iterations,
                                  "launches": best.kernel_launches, "rays": best.rays, "mpaths_s": best.paths / best.render_ms / 1e3}), flush=True)


if __name__ == "__main__":
    main()
