"""The drop-in boundary: libptcore.so loads without a GPU, exports every symbol include/ptcore.h declares, keeps the
struct layouts the header promises, reports errors as codes (never aborts), and REFUSES to compute without a CUDA
device — there is no CPU fallback behind the C ABI."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    src = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, src)))


def test_every_declared_symbol_is_exported(pt):
    core, host = pt.core(), pt.host()
    names = _declared("ptcore.h", "ptc")
    assert len(names) >= 20
    for n in names:
        assert hasattr(core, n), n
    for n in _declared("pthost.h", "pth"):
        assert hasattr(host, n), n
    assert core.ptc_abi_version() == 1


def test_struct_layouts_match_the_header(pt, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ptcore.h"\n#include "pthost.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(ptc_material),sizeof(ptc_camera),sizeof(ptc_render_settings),sizeof(ptc_hit),sizeof(ptc_stats),"
                   "sizeof(ptc_mesh_info),sizeof(pth_object),offsetof(ptc_render_settings,seed),offsetof(ptc_stats,nodes_visited));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])  # the header is plain C
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(pt.Material), C.sizeof(pt.Camera), C.sizeof(pt.RenderSettings), C.sizeof(pt.Hit), C.sizeof(pt.Stats),
            C.sizeof(pt.MeshInfo), C.sizeof(pt.HostObject), pt.RenderSettings.seed.offset, pt.Stats.nodes_visited.offset]
    assert got == want


def test_product_does_not_depend_on_the_oracle():
    # nothing under the package may reference oracle/ or the hostsim harness
    pkg = os.path.join(ROOT, "raytracer-rust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".so", ".pyc")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            assert "liboracle" not in text and "orc_" not in text and "hostsim.cpp" not in text, os.path.join(dirpath, f)
    ldd = subprocess.check_output(["ldd", os.path.join(pkg, "libptcore.so")], text=True)
    assert "oracle" not in ldd and "hostsim" not in ldd
    sass = subprocess.run(["cuobjdump", "-lelf", os.path.join(pkg, "libptcore.so")], capture_output=True, text=True)
    if sass.returncode == 0:
        assert "sm_100a" in sass.stdout  # the kernels are built for B200, nothing else


def test_error_codes_and_call_order(pt):
    core = pt.core()
    s = pt.Scene()
    m = s.add_material(pt.lambertian((1, 1, 1)))
    s.add_sphere((0, 0, 0), 1.0, m)
    cs = s.to_core()
    st = pt.RenderSettings(width=8, height=8, spp=1, max_depth=2)
    cam = pt.camera_new((0, 0, 3), (0, 0, 0), (0, 1, 0), 40.0, 1.0)
    with pytest.raises(pt.PtcError) as e:
        cs.render(cam, st)  # render before commit
    assert e.value.code == pt.PTC_E_STATE
    # bad material index / null mesh: PTC_E_INVALID, message available
    h = core.ptc_scene_create()
    c3 = (C.c_float * 3)(0, 0, 0)
    assert core.ptc_scene_add_sphere(h, c3, 1.0, 5) == pt.PTC_E_INVALID
    assert b"material" in core.ptc_last_error()
    ident = (C.c_float * 16)(*[1 if i % 5 == 0 else 0 for i in range(16)])
    assert core.ptc_scene_add_material(h, C.byref(pt.lambertian((1, 1, 1)))) == 0
    assert core.ptc_scene_add_mesh(h, None, 0, ident, ident, 0) == pt.PTC_E_INVALID  # mesh_object.rs:30-36
    bad = pt.Material(type=99)
    assert core.ptc_scene_add_material(h, C.byref(bad)) == pt.PTC_E_INVALID
    # a sphere the reference cannot intersect without panicking (Vec3 / radius with |radius| < 1e-4, vec3.rs:117-122) is
    # refused at the boundary instead: an error code, never an unwind or a silent division
    c = (C.c_float * 3)(0, 0, 0)
    assert core.ptc_scene_add_sphere(h, c, 5e-5, 0) == pt.PTC_E_INVALID and b"panic in the reference" in core.ptc_last_error()
    assert core.ptc_scene_add_sphere(h, c, -5e-5, 0) == pt.PTC_E_INVALID
    assert core.ptc_scene_add_sphere(h, c, float("nan"), 0) == pt.PTC_E_INVALID
    assert core.ptc_scene_add_sphere(h, c, 1e-4, 0) >= 0
    core.ptc_scene_destroy(h)


def test_no_cpu_fallback(pt):
    if pt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    s = pt.Scene()
    m = s.add_material(pt.lambertian((1, 1, 1)))
    s.add_sphere((0, 0, 0), 1.0, m)
    cs = s.to_core()
    with pytest.raises(pt.PtcError) as e:
        cs.commit(0)
    assert e.value.code == pt.PTC_E_CUDA and "no CPU fallback" in str(e.value)
    out = (C.c_uint32 * 64)()
    assert pt.host().pth_render_scene(s._h, 0, out, None) == pt.PTC_E_CUDA


def test_host_build_needs_no_device(pt, scenes_dir):
    # the host half of commit (reference-BVH dead mask -> SAH -> 8-wide quantised BVH) is plain C++
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "semesterbild.json"))
    cs = s.to_core().build()
    info, dead, order = cs.mesh_info(3)
    assert (info.ref_nodes, info.ref_leaves, info.ref_depth) == (3351, 1676, 11)
    assert info.live_triangles == 2766 and int(dead.sum()) == 1982
    assert info.wide_depth <= 8 and info.wide_nodes * 80 == info.node_bytes
    with pytest.raises(pt.PtcError):
        cs.mesh_info(0)  # a cube is not a mesh
