#!/usr/bin/env python3
"""SIMT cost model of the wavefront stages on the CPU (tests/hostsim/wfsim.cpp): ranks scheduling designs before a GPU
run.  TOOLING ONLY — the numbers that count are measured on the B200 (profiles/).

  python tools/simt_model.py [c2|c3|c4|c5s|c1] [--iterations 12] [--cap 14080]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ptload  # noqa: E402

pt = ptload.load()
from bindings import SimScene, sim_lib  # noqa: E402


class Params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32), ("iterations", C.c_int32),
                ("cap", C.c_int32), ("n_policies", C.c_int32), ("seed", C.c_uint64), ("policy", (C.c_int32 * 5) * 8)]


def scene_for(name):
    SC = os.path.join(ROOT, "scenes")
    if name == "c2":
        return pt.load_scene_from_json(os.path.join(SC, "semesterbild.json"))
    if name == "c1":
        s = pt.load_scene_from_json(os.path.join(SC, "cornell-box", "scene.json"))
        s.set_settings(256, 256, 16, 8)
        return s
    if name == "c4":
        return pt.load_scene_from_json(os.path.join(SC, "veach-mis", "scene.json"))
    if name == "c3":
        import bench_configs
        return bench_configs.teapot_scene()
    if name == "c5s":  # C5's scene with a 250x250-cell field (125 k triangles): same structure, CPU-sized
        s = pt.synthetic_scene(cells=250)
        return s
    raise SystemExit("unknown scene " + name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", nargs="?", default="c2")
    ap.add_argument("--iterations", type=int, default=12)
    ap.add_argument("--cap", type=int, default=14080)
    ap.add_argument("--policies", default="8,0,0,0,2;8,1,8,8,2;8,1,12,12,2;8,1,16,16,2;8,2,8,8,2;4,0,0,0,2;12,0,0,0,2;8,1,12,12,4")
    a = ap.parse_args()
    scene = scene_for(a.scene)
    w, h, spp, depth = scene.settings
    sim = SimScene(scene)
    L = sim_lib()
    L.sim_wavefront_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int]
    p = Params(width=w, height=h, max_depth=depth, iterations=a.iterations, cap=a.cap, seed=0)
    pols = [[int(x) for x in q.split(",")] for q in a.policies.split(";")]
    p.n_policies = len(pols)
    for i, q in enumerate(pols):
        for j in range(5):
            p.policy[i][j] = q[j]
    out = (C.c_double * 128)()
    cam = scene.camera
    n = L.sim_wavefront_model(sim.h, C.addressof(cam), C.addressof(p), out, 128)
    assert n > 0, n
    rays, tasks, pairs, npairs, masksort, now, lanecull, blockc, tests, tpass, hits, wobj, wobj_pass, shade_eff, s256, y256, s2048, y2048 = out[:18]
    print(f"scene {a.scene}: {w}x{h} depth {depth}, {a.iterations} iterations of a {a.cap}-slot segment: {rays:.0f} rays, "
          f"{tasks:.0f} mesh tasks ({tasks / rays:.3f}/ray)")
    print(f"extend pre/post scans: exact tests {tests:.0f} ({tests / rays:.2f}/ray), pass world-box cull {tpass / tests:.3f}, "
          f"accepted hits {hits / tests:.3f}; warps x objects with >=1 passing lane {wobj_pass / wobj:.3f}")
    print(f"  model warp-instr/ray: v1 {now / rays * 32:.0f}  per-lane cull {lanecull / rays * 32:.0f} ({lanecull / now:.2f}x)  "
          f"block compaction {blockc / rays * 32:.0f} ({blockc / now:.2f}x)  sort by cull mask {masksort / rays * 32:.0f} ({masksort / now:.2f}x)  dense pairs {pairs / rays * 32:.0f} ({pairs / now:.2f}x, {npairs / rays:.2f} pairs/ray)")
    print(f"shade: lane efficiency after the class sort {shade_eff:.3f}; model warp-instr/ray x32: 256-window sum {s256 / rays * 32:.0f}, "
          f"barrier-synchronous {y256 / rays * 32:.0f}; 2048-window sum {s2048 / rays * 32:.0f}, synchronous {y2048 / rays * 32:.0f}")
    base = None
    for i, q in enumerate(pols):
        cost, ideal, turns, nexec, texec, nl, tl, refills = out[18 + 8 * i: 26 + 8 * i]
        if cost == 0:
            continue
        base = base or cost
        print(f"traverse policy refill={q[0]} mode={q[1]} thr=({q[2]},{q[3]}) groups={q[4]}: cost {cost / base:.3f} of first, "
              f"efficiency {ideal / cost:.3f}, node lanes {nl / max(nexec, 1):.1f}/32 x {nexec:.0f}, tri lanes {tl / max(texec, 1):.1f}/32 x {texec:.0f}, "
              f"nodes/task {nl / tasks:.2f} tris/task {tl / tasks:.2f}")
    if tasks == 0:
        return
    k = 18 + 8 * len(pols)
    hn, ht, hits = list(out[k:k + 16]), list(out[k + 16:k + 32]), out[k + 32]
    print("node steps per task histogram (last bin = 15+):", " ".join(f"{v / tasks:.3f}" for v in hn))
    print("tri tests per task histogram:", " ".join(f"{v / tasks:.3f}" for v in ht))
    print(f"tasks that end with a triangle hit: {hits / tasks:.3f}")


if __name__ == "__main__":
    main()
