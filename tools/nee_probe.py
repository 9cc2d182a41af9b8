#!/usr/bin/env python3
"""What PTC_FLAG_NEE buys on veach-mis (BASELINE config C4, native 1280x720, depth 16) and the Cornell box: error against
a converged render at equal spp and the render time of both integrators.  One JSON line per scene."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptload  # noqa: E402

pt = ptload.load()
from raytracer_rust_b200 import workloads  # noqa: E402


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


for cfg, spp, ref_spp in (("C4", 64, 8192), ("C1", 64, 16384)):
    label, s = workloads.workload(cfg)
    w, h, _, depth = s.settings
    cs = s.to_core().commit(0)
    ref, _ = cs.render(s.camera, s.render_settings(spp=ref_spp, seed=1, flags=pt.FLAG_NEE))
    ref2, _ = cs.render(s.camera, s.render_settings(spp=ref_spp, seed=9))
    row = {"config": label, "spp": spp, "reference": f"NEE render at {ref_spp} spp", "relmse_plain_vs_nee_references": relmse(ref2, ref)}
    for name, flags in (("plain", 0), ("nee", pt.FLAG_NEE)):
        cs.render(s.camera, s.render_settings(spp=spp, seed=2, flags=flags))
        img, st = cs.render(s.camera, s.render_settings(spp=spp, seed=2, flags=flags))
        row[name] = {"relmse": relmse(img, ref), "render_ms": st.render_ms, "rays": st.rays, "mpaths_s": st.paths / st.render_ms / 1e3}
    row["error_ratio_plain_over_nee"] = row["plain"]["relmse"] / row["nee"]["relmse"]
    row["time_ratio_nee_over_plain"] = row["nee"]["render_ms"] / row["plain"]["render_ms"]
    row["efficiency_gain"] = row["error_ratio_plain_over_nee"] / row["time_ratio_nee_over_plain"]
    print(json.dumps(row), flush=True)
