#!/usr/bin/env python3
"""Probe: C2 at its native settings through ptc_render_accumulate on torch's default stream against ptc_render_u32 on the
library's own stream, for several pool sizes (0 = library default)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
from raytracer_rust_b200 import workloads  # noqa: E402

label, scene = workloads.workload("C2")
cs = scene.to_core().commit(0)
w, h, spp, depth = scene.settings
accum = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda:0")
stream = torch.cuda.current_stream()
for pool in (48 << 20, 128 << 20, 0, 0):
    st = scene.render_settings(spp=spp, seed=0, pool_paths=pool)
    for mode in ("accumulate on torch stream", "render_u32 on own stream"):
        best_ev, best_wall, it = 1e9, 1e9, 0
        for _ in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            if mode.startswith("acc"):
                accum.zero_()
                s = cs.render_accumulate(scene.camera, st, accum.data_ptr(), stream.cuda_stream)
            else:
                _, s = cs.render_u32(scene.camera, st)
            e1.record(stream)
            torch.cuda.synchronize()
            best_wall = min(best_wall, (time.perf_counter() - t0) * 1e3)
            best_ev = min(best_ev, e0.elapsed_time(e1))
            it = s.iterations
        print(json.dumps({"pool": pool >> 20, "mode": mode, "event_ms": round(best_ev, 2), "wall_ms": round(best_wall, 2), "render_ms": round(s.render_ms, 2),
                          "iterations": it, "pool_slots_Mi": cs.last_pool_slots() >> 20}), flush=True)
