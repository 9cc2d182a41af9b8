"""A plain-C program (tests/c_abi_driver.c, gcc, include/ptcore.h only) drives libptcore.so the way the reference's Rust
host would after the patch of INTEGRATION.md section 4: ptc_scene_add_* in object_list order, commit, ptc_render_u32.
The closest thing to the `extern "C"` caller this image allows (no rustc)."""
import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracer-rust_b200")


def build_driver(tmp_path):
    exe = str(tmp_path / "c_abi_driver")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_driver.c"), "-o", exe, "-L", PKG, "-lptcore", f"-Wl,-rpath,{PKG}"])
    return exe


def dump_scene(pt, scene, st, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 0x31435450))
        f.write(bytes(scene.camera))
        f.write(struct.pack("<iiiiQ", st.width, st.height, st.spp, st.max_depth, st.seed))
        mats = scene.materials
        f.write(struct.pack("<i", len(mats)))
        for m in mats:
            f.write(bytes(m))
        objs = scene.objects
        f.write(struct.pack("<i", len(objs)))
        fl = lambda *a: np.concatenate([np.asarray(x, np.float32).ravel() for x in a]).tobytes()  # noqa: E731
        for o in objs:
            f.write(struct.pack("<ii", o.type, o.material))
            if o.type == pt.OBJ_SPHERE:
                f.write(fl(list(o.center), [o.radius]))
            elif o.type == pt.OBJ_PLANE:
                f.write(fl(list(o.p1), list(o.normal)))
            elif o.type == pt.OBJ_QUAD:
                f.write(fl(list(o.base), list(o.edge0), list(o.edge1), list(o.normal), [o.d, o.inv_edge0_len_sq, o.inv_edge1_len_sq]))
            elif o.type == pt.OBJ_CUBE:
                f.write(fl(list(o.o2w), list(o.w2o)))
            else:
                t = scene.mesh(o.mesh)
                f.write(fl(list(o.o2w), list(o.w2o)))
                f.write(struct.pack("<q", len(t)))
                f.write(np.ascontiguousarray(t, np.float32).tobytes())
        sky = scene.sky
        if sky is None:
            f.write(struct.pack("<ii", 0, 0))
        else:
            f.write(struct.pack("<ii", sky.shape[1], sky.shape[0]))
            f.write(np.ascontiguousarray(sky, np.float32).tobytes())


def test_c_driver_builds_links_and_has_no_cpu_fallback(pt, scenes_dir, tmp_path):
    """Without a GPU: the C program compiles against include/ptcore.h with -Werror, links libptcore.so, replays the Cornell
    box through the add_* calls and — on a box without a CUDA device — gets PTC_E_CUDA from ptc_scene_commit (exit 2)."""
    exe = build_driver(tmp_path)
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "cornell-box", "scene.json"))
    st = s.render_settings(width=32, height=32, spp=2, max_depth=4, seed=1)
    dump_scene(pt, s, st, tmp_path / "scene.bin")
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "out.u32")], capture_output=True, text=True)
    if pt.device_count() == 0:
        assert r.returncode == 2 and "ptc_scene_commit" in r.stderr, (r.returncode, r.stderr)
    else:
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,spp,depth", [("cornell-box/scene.json", 128, 128, 8, 8), ("semesterbild.json", 200, 150, 4, 30)])
def test_c_driver_renders_what_the_python_route_renders(pt, scenes_dir, tmp_path, name, w, h, spp, depth):
    exe = build_driver(tmp_path)
    s = pt.load_scene_from_json(os.path.join(scenes_dir, name))
    st = s.render_settings(width=w, height=h, spp=spp, max_depth=depth, seed=5)
    dump_scene(pt, s, st, tmp_path / "scene.bin")
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "out.u32")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stats = json.loads(r.stdout)
    got = np.fromfile(tmp_path / "out.u32", np.uint32)
    want, ws = s.to_core().commit(0).render_u32(s.camera, st)
    assert stats["paths"] == ws.paths == w * h * spp and stats["rays"] == ws.rays
    assert np.array_equal(got, want)  # same seed, same paths, fixed-point film: the same Vec<u32> bit for bit
