#!/usr/bin/env python3
"""How sensitive is the flat-node (dead-triangle) set of the reference BVH to the order of EQUAL centroid keys?

BVHNode::new sorts every node's triangles with `sort_unstable_by` on one centroid coordinate
(/root/reference/src/acceleration/bvh.rs:45-53); Rust leaves the order of equal keys unspecified, it differs between std
versions (pdqsort up to 1.80, ipnsort since) and neither can be run here.  Which triangles share a zero-extent node —
and are therefore invisible (aabb.rs:40) — depends on it.  Oracle and product both use a stable sort; this tool measures
what the choice is worth on the default scene's mesh (RayTracingText.obj, 4,748 triangles, 1,982 dead under the stable
order), with the oracle only (CPU):

  * dead-mask differences against the stable order for `reverse` and for K random tie orders,
  * for each, an 8-spp 800x600 oracle render against the reference's own docs/semesterbild.png (8x8 block means,
    tests/golden/): mean block difference, blocks off by more than 10 levels, and which of those blocks are shared.

  python tools/tie_order_study.py [--random 6] [--spp 8] > profiles/r2_tie_order_study.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ptload  # noqa: E402

pt = ptload.load()
from bindings import RNG_CHACHA, OracleScene, oracle_resolve  # noqa: E402


def blocks(u32):
    u = np.asarray(u32, np.uint32).reshape(600, 800)
    rgb = np.stack([(u >> 16) & 255, (u >> 8) & 255, u & 255], -1).astype(np.float32)
    return rgb.reshape(75, 8, 100, 8, 3).mean(axis=(1, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--random", type=int, default=6)
    ap.add_argument("--spp", type=int, default=8)
    a = ap.parse_args()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "semesterbild_ref_blocks.npy"))
    scene = pt.load_scene_from_json(os.path.join(ROOT, "scenes", "semesterbild.json"))
    mesh_obj = [i for i, o in enumerate(scene.objects) if o.type == pt.OBJ_MESH][0]
    n_tris = len(scene.mesh(scene.objects[mesh_obj].mesh))
    modes = ["stable", "reverse"] + [f"random:{k + 1}" for k in range(a.random)]
    rows, base_dead, base_off = [], None, None
    for mode in modes:
        os.environ["PTC_REF_TIE"] = mode
        orc = OracleScene(scene)  # the BVH is built here, under the tie order just selected
        nodes, leaves, depth, dead, order = orc.mesh_bvh_info(mesh_obj, n_tris)
        # the product's host builder under the same switch must give the same mask and DFS order
        info, pdead, porder = scene.to_core().build().mesh_info(mesh_obj)
        assert (pdead == dead).all() and (porder == order).all(), mode
        img, _ = orc.render(scene.camera, 800, 600, a.spp, 30, rng_mode=RNG_CHACHA)
        d = np.abs(blocks(oracle_resolve(img)) - gold)
        off = d.max(-1) > 10
        if base_dead is None:
            base_dead, base_off = dead.copy(), off.copy()
        rows.append({"tie_order": mode, "dead_triangles": int(dead.sum()), "dead_changed_vs_stable": int((dead != base_dead).sum()),
                     "newly_dead": int(((dead == 1) & (base_dead == 0)).sum()), "newly_live": int(((dead == 0) & (base_dead == 1)).sum()),
                     "ref_nodes": nodes, "ref_leaves": leaves, "ref_depth": depth,
                     "block_mean_abs_diff_levels": float(d.mean()), "blocks_off_by_more_than_10": int(off.sum()),
                     "off_blocks_shared_with_stable": int((off & base_off).sum()),
                     "off_blocks": [[int(y), int(x)] for y, x in zip(*np.nonzero(off))]})
        print(json.dumps({k: v for k, v in rows[-1].items() if k != "off_blocks"}), file=sys.stderr, flush=True)
    os.environ.pop("PTC_REF_TIE", None)
    print(json.dumps({"mesh": "scenes/RayTracingText.obj", "triangles": n_tris, "spp": a.spp, "golden": "docs/semesterbild.png of the reference, 8x8 block means",
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
