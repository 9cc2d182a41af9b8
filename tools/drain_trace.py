#!/usr/bin/env python3
"""Per-iteration stage times of one render (PTC_FLAG_TIMING: events between the launches, no programmatic dependent
launch — the times include one event each and are upper bounds), to see what the latency floor of a drain iteration is
made of.  One JSON line: for every iteration the microseconds of pre / traverse / post / shade.

  python tools/drain_trace.py C2 [share]        (pool = the library's default)"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
from raytracer_rust_b200 import workloads  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
share = int(sys.argv[2]) if len(sys.argv) > 2 else 8
label, scene = workloads.workload(cfg)
cs = scene.to_core().commit(0)
w, h, spp, depth = scene.settings
st = scene.render_settings(spp=spp, sample_begin=0, sample_end=max(1, spp // share), seed=0, pool_paths=0)
for _ in range(3):
    _, plain = cs.render_u32(scene.camera, st)
st.flags = pt.FLAG_TIMING
path = os.path.join(tempfile.mkdtemp(), "stages.txt")
os.environ["PTC_TRACE_STAGES"] = path
_, s = cs.render_u32(scene.camera, st)
del os.environ["PTC_TRACE_STAGES"]
iters, cur = [], None
for line in open(path):
    stage, ms = line.split()
    stage, us = int(stage), float(ms) * 1e3
    if stage == 0:  # a pre opens an iteration (the very first mark is the initial fill's shade)
        cur = [0.0, 0.0, 0.0, 0.0]
        iters.append(cur)
    if cur is not None:
        cur[stage] += us
print(json.dumps({"config": label, "share": share, "render_ms_plain": plain.render_ms, "render_ms_timed": s.render_ms, "iterations": s.iterations,
                  "us_pre_traverse_post_shade": [[round(v, 1) for v in it] for it in iters]}))
