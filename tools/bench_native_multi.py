#!/usr/bin/env python3
"""The in-process multi-GPU path of the C ABI (ptc_multi_*: one host thread per GPU, one NCCL reduce of the film, no
torch, no torchrun) on 1..N GPUs of this node: the route a single-process host such as the reference's Rust binary takes.

  python tools/bench_native_multi.py --config c5 --spp 64 --gpus 1,2,4,8 [--shard tiles|samples] [--steps 3]

The job is FIXED (the config's frame at --spp samples): strong scaling.  Times are host wall clock around the whole
call (renders on all devices + reduce + film resolve on device 0 + copy of the packed image to the host), i.e. end to
end with HOST buffers, like `render_scene` returning its Vec<u32>.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ptload  # noqa: E402

pt = ptload.load()
from bench_multi import scene_for  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--shard", default="tiles", choices=["tiles", "samples"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--pool", type=int, default=3 << 22)
    a = ap.parse_args()
    scene, label = scene_for(pt, a.config)
    cs = scene.to_core().commit(0)
    st = scene.render_settings(spp=a.spp, seed=0, pool_paths=a.pool)
    shard = pt.SHARD_TILES if a.shard == "tiles" else pt.SHARD_SAMPLES
    have = pt.device_count()
    for n in [int(x) for x in a.gpus.split(",")]:
        if n > have:
            continue
        m = cs.multi(list(range(n)))
        m.render_u32(scene.camera, st, shard)  # warm-up: pools, NCCL channels
        best, tot, stats = None, 0.0, None
        for _ in range(a.steps):
            t0 = time.perf_counter()
            _, stats = m.render_u32(scene.camera, st, shard)  # render_scene's Vec<u32>: the film is resolved on device 0
            dt = (time.perf_counter() - t0) * 1e3
            tot += dt
            best = dt if best is None else min(best, dt)
        ms = tot / a.steps
        print(json.dumps({"config": label, "path": "ptc_multi_render (in-process, NCCL reduce)", "n_gpus": n, "sharding": a.shard,
                          "spp": a.spp, "steps": a.steps, "ms_per_step": ms, "best_ms": best, "mpaths_per_s": stats.paths / ms / 1e3,
                          "mrays_per_s": stats.rays / ms / 1e3, "film_reduce_bytes": st.width * st.height * 12 if n > 1 else 0,
                          "d2h_bytes": st.width * st.height * 4}), flush=True)
        del m


if __name__ == "__main__":
    main()
