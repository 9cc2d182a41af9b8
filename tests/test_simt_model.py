"""The SIMT cost model (tests/hostsim/wfsim.cpp, tools/simt_model.py) is tooling, but its replay of the wavefront has
to stay in step with the kernels: the conservative cull it evaluates must never reject an accepted hit (the C side
returns -7 if it does), and the traversal it replays must visit what the straight-line traversal visits."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("name", ["c2", "c4"])
def test_wavefront_model_runs_and_is_consistent(pt, name):
    import simt_model
    from bindings import SimScene, sim_lib

    scene = simt_model.scene_for(name)
    w, h, spp, depth = scene.settings
    sim = SimScene(scene)
    L = sim_lib()
    L.sim_wavefront_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int]
    p = simt_model.Params(width=w, height=h, max_depth=depth, iterations=3, cap=2048, seed=0, n_policies=2)
    for i, q in enumerate([[8, 0, 0, 0, 2], [4, 1, 8, 8, 2]]):
        for j in range(5):
            p.policy[i][j] = q[j]
    out = (C.c_double * 128)()
    cam = scene.camera
    n = L.sim_wavefront_model(sim.h, C.addressof(cam), C.addressof(p), out, 128)
    assert n > 0, "model failed (a negative value means the world-box cull rejected an accepted hit)"
    rays, tasks = out[0], out[1]
    assert rays >= 3 * 2048 * 0.5
    k = 18
    cost0, ideal0, turns0, nexec0, texec0, nl0, tl0, _ = out[k:k + 8]
    cost1, ideal1, turns1, nexec1, texec1, nl1, tl1, _ = out[k + 8:k + 16]
    if name == "c2":
        assert tasks > 0 and 0.2 < ideal0 / cost0 < 1.0
        # scheduling policies change the order of the steps, never the set of nodes a ray must visit by more than the
        # effect of a later-shrinking t_max
        assert abs(nl1 - nl0) / nl0 < 0.1 and abs(tl1 - tl0) / tl0 < 0.2
    else:
        assert tasks == 0
