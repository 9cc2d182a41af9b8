"""Host stand-in (libpthost.so) against the reference loader's documented behaviour (src/tungsten/parser.rs:245-815,
SURVEY.md Appendix A.8).  Pure host code: no GPU."""
import json
import math
import os

import numpy as np
import pytest


def test_semesterbild_as_shipped(pt, scenes_dir):
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "semesterbild.json"))
    assert s.settings == (800, 600, 256, 30)
    objs = s.objects
    assert [o.type for o in objs] == [pt.OBJ_CUBE] * 3 + [pt.OBJ_MESH, pt.OBJ_SPHERE]
    mats = s.materials
    assert mats[objs[0].material].type == pt.MAT_ROUGH_CONDUCTOR
    al = mats[objs[0].material]
    assert list(al.eta) == pytest.approx([1.36, 0.965, 0.62]) and list(al.k) == pytest.approx([7.57, 6.69, 5.44])
    assert al.distribution == pt.DIST_GGX and al.roughness == pytest.approx(0.1)
    assert mats[objs[4].material].type == pt.MAT_DIELECTRIC and mats[objs[4].material].ior == pytest.approx(1.52)
    assert objs[4].radius == 50.0
    assert s.mesh(0).shape == (4748, 12)
    assert s.sky is None


def test_cornell_box_geometry_pins_euler_order(pt, scenes_dir):
    # A.8: with T*Ry*Rx*Rz*S the BackWall quad (pos (0,1,-1), rot (0,90,90), scale (2,4,2)) lies in z = -1 and spans
    # x in [-1,1], y in [0,2]; RightWall in x = 1
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "cornell-box", "scene.json"))
    assert s.settings == (1024, 1024, 64, 64)
    objs = s.objects
    assert [o.type for o in objs] == [pt.OBJ_QUAD] * 5 + [pt.OBJ_CUBE] * 2 + [pt.OBJ_QUAD]
    back = objs[2]
    corners = np.array([np.array(back.base), np.array(back.base) + back.edge0, np.array(back.base) + back.edge1,
                        np.array(back.base) + back.edge0 + back.edge1])
    assert np.allclose(corners[:, 2], -1.0, atol=1e-5)
    assert np.allclose(sorted(set(np.round(corners[:, 0], 4))), [-1, 1]) and np.allclose(sorted(set(np.round(corners[:, 1], 4))), [0, 2])
    right = objs[3]
    assert abs(abs(right.normal[0]) - 1) < 1e-5 and abs(right.base[0] - 1) < 1e-5
    # the light is a quad whose `emission` replaces its bsdf (parser.rs:707-711)
    light = s.materials[objs[7].material]
    assert light.type == pt.MAT_EMISSIVE and list(light.albedo) == pytest.approx([17, 12, 4])
    # "null" bsdf -> black Lambertian (parser.rs:357-359)
    assert any(m.type == pt.MAT_LAMBERT and list(m.albedo) == [0, 0, 0] for m in s.materials)


def test_veach_mis_materials(pt, scenes_dir):
    s = pt.load_scene_from_json(os.path.join(scenes_dir, "veach-mis", "scene.json"))
    assert s.settings == (1280, 720, 1024, 16)
    mats, objs = s.materials, s.objects
    rough = sorted(round(mats[o.material].roughness, 3) for o in objs if o.type == pt.OBJ_CUBE)
    assert rough == [0.01, 0.05, 0.1, 0.25]
    assert all(mats[o.material].distribution == pt.DIST_BECKMANN for o in objs if o.type == pt.OBJ_CUBE)
    cu = mats[[o for o in objs if o.type == pt.OBJ_CUBE][0].material]
    assert list(cu.eta) == pytest.approx([0.2, 1.09, 1.42])


def _write(tmp_path, doc):
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(doc))
    return str(p)


CAM = {"transform": {"position": [0, 0, 5], "look_at": {"x": 0, "y": 0, "z": 0}, "up": [0, 1, 0]}, "fov": 40}


def test_defaults_and_both_vec3_forms(pt, tmp_path):
    s = pt.load_scene_from_json(_write(tmp_path, {"camera": CAM, "primitives": []}))
    assert s.settings == (800, 600, 16, 10)  # parser.rs:255-258
    cam = s.camera
    assert cam.half_height == pytest.approx(math.tan(math.radians(20)), rel=1e-6)
    assert cam.half_width == pytest.approx(cam.half_height * 800 / 600, rel=1e-6)  # aspect defaults to W/H
    assert list(cam.forward) == pytest.approx([0, 0, -1])


def test_resolution_variants_and_overrides(pt, tmp_path):
    doc = {"camera": dict(CAM, resolution=[320, 200], aspect=2.0), "primitives": [], "renderer": {"spp": 7},
           "integrator": {"max_bounces": 3}}
    s = pt.load_scene_from_json(_write(tmp_path, doc))
    assert s.settings == (320, 200, 7, 3)
    assert s.camera.half_width == pytest.approx(s.camera.half_height * 2.0)
    s = pt.load_scene_from_json(_write(tmp_path, {"camera": dict(CAM, resolution=64), "primitives": []}))
    assert s.settings[:2] == (64, 64)


def test_unknown_primitive_type_fails_the_whole_file(pt, tmp_path):
    # serde tag enum: `infinite_sphere` is not a variant (parser.rs:135-165) -> the shipped teapot scene cannot load
    doc = {"camera": CAM, "primitives": [{"type": "infinite_sphere", "transform": {}}]}
    with pytest.raises(RuntimeError, match="unknown variant"):
        pt.load_scene_from_json(_write(tmp_path, doc))


def test_missing_required_fields(pt, tmp_path):
    with pytest.raises(RuntimeError, match="camera"):
        pt.load_scene_from_json(_write(tmp_path, {"primitives": []}))
    with pytest.raises(RuntimeError, match="transform"):
        pt.load_scene_from_json(_write(tmp_path, {"camera": CAM, "primitives": [{"type": "cube", "bsdf": "x"}]}))
    with pytest.raises(RuntimeError, match="bsdf"):
        pt.load_scene_from_json(_write(tmp_path, {"camera": CAM, "primitives": [{"type": "cube", "transform": {}}]}))
    with pytest.raises(RuntimeError):
        pt.load_scene_from_json(str(tmp_path / "does_not_exist.json"))


def test_bsdf_fallbacks(pt, tmp_path):
    doc = {"camera": CAM, "bsdfs": [
        {"name": "c", "type": "conductor"},                      # unsupported type -> skipped
        {"name": "p", "type": "plastic"},                        # defaults: albedo .8, ior 1.5
        {"name": "g", "type": "lambert", "albedo": 0.25},        # grayscale
        {"name": "k", "type": "lambert", "albedo": {"type": "checker", "on_color": [1, 1, 1], "off_color": [0, 0, 0], "res_u": 4}},
        {"name": "r", "type": "rough_conductor", "roughness": 0.001, "material": "AU", "distribution": "Beckmann"},
    ], "primitives": [
        {"type": "cube", "transform": {}, "bsdf": "c"}, {"type": "cube", "transform": {}, "bsdf": "p"},
        {"type": "cube", "transform": {}, "bsdf": "g"}, {"type": "cube", "transform": {}, "bsdf": "k"},
        {"type": "cube", "transform": {}, "bsdf": "r"}, {"type": "sphere", "transform": {"scale": [3, 3, 3]}, "bsdf": "nope"},
    ]}
    s = pt.load_scene_from_json(_write(tmp_path, doc))
    mats, objs = s.materials, s.objects
    m = [mats[o.material] for o in objs]
    assert m[0].type == pt.MAT_LAMBERT and list(m[0].albedo) == [1, 0, 1]           # magenta
    assert m[1].type == pt.MAT_PLASTIC and list(m[1].albedo) == pytest.approx([0.8] * 3) and m[1].ior == 1.5
    assert list(m[2].albedo) == [0.25] * 3
    assert m[3].type == pt.MAT_LAMBERT_CHECKER and m[3].inv_scale == pytest.approx(0.25)
    assert m[4].roughness == pytest.approx(0.01) and m[4].distribution == pt.DIST_BECKMANN  # max(0.01), case-insensitive
    assert list(m[4].eta) == pytest.approx([0.17, 0.35, 1.5])
    assert list(m[5].albedo) == [1, 0, 1] and objs[5].radius == 3.0                   # radius from scale.x


def test_inline_plane_material_and_quad_string_emission(pt, tmp_path):
    doc = {"camera": CAM, "bsdfs": [{"name": "w", "type": "lambert", "albedo": [1, 1, 1]}], "primitives": [
        {"type": "plane", "point": [0, -1, 0], "normal": [0, 2, 0], "material": {"Metal": {"albedo": [0.9, 0.8, 0.7], "fuzz": 3.0}}},
        {"type": "quad", "transform": {}, "bsdf": "w", "emission": "textures/light.png"},
    ]}
    s = pt.load_scene_from_json(_write(tmp_path, doc))
    plane, quad = s.objects
    assert plane.type == pt.OBJ_PLANE and list(plane.normal) == pytest.approx([0, 1, 0])  # Plane::new normalises
    metal = s.materials[plane.material]
    assert metal.type == pt.MAT_METAL and metal.fuzz == 1.0                              # Metal::new clamps fuzz
    assert list(s.materials[quad.material].albedo) == [5, 5, 5]                          # parser.rs:712-716


def test_quad_canonical_frame(pt):
    # unit quad in the local XZ plane, normal = e0 x e1 = local -Y (quad.rs:31-52)
    s = pt.Scene()
    m = s.add_material(pt.lambertian((1, 1, 1)))
    s.add_quad(m)
    q = s.objects[0]
    assert list(q.base) == [-0.5, 0, -0.5] and list(q.edge0) == [1, 0, 0] and list(q.edge1) == [0, 0, 1]
    assert list(q.normal) == [0, -1, 0] and q.inv_edge0_len_sq == 1.0 and q.d == 0.0


def test_transform_is_t_ry_rx_rz_s_and_inverse(pt):
    o2w, w2o = pt.transform((2, 3, 4), (10, 20, 30), (5, 6, 7))
    M, Mi = o2w.reshape(4, 4).T.astype(np.float64), w2o.reshape(4, 4).T.astype(np.float64)
    rx, ry, rz = np.radians([10, 20, 30])
    Rx = np.array([[1, 0, 0], [0, math.cos(rx), -math.sin(rx)], [0, math.sin(rx), math.cos(rx)]])
    Ry = np.array([[math.cos(ry), 0, math.sin(ry)], [0, 1, 0], [-math.sin(ry), 0, math.cos(ry)]])
    Rz = np.array([[math.cos(rz), -math.sin(rz), 0], [math.sin(rz), math.cos(rz), 0], [0, 0, 1]])
    want = np.eye(4)
    want[:3, :3] = Ry @ Rx @ Rz @ np.diag([2, 3, 4])
    want[:3, 3] = [5, 6, 7]
    assert np.allclose(M, want, atol=1e-5)
    assert np.allclose(M @ Mi, np.eye(4), atol=1e-5)


def test_obj_loader_fan_triangulation_and_degenerate_filter(pt, tmp_path):
    p = tmp_path / "m.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 2 2 2\nvn 0 0 1\n"
                 "f 1/1/1 2/1/1 3/1/1 4/1/1\n"      # quad -> (0,1,2), (0,2,3)
                 "f -1 -1 -1\n"                      # degenerate -> filtered (mesh_object.rs:128-134)
                 "l 1 2\n")
    s = pt.Scene()
    m = s.add_material(pt.lambertian((1, 1, 1)))
    s.add_obj(str(p), m)
    t = s.mesh(0)
    assert t.shape == (2, 12)
    assert t[0, :9].tolist() == [0, 0, 0, 1, 0, 0, 1, 1, 0] and t[1, :9].tolist() == [0, 0, 0, 1, 1, 0, 0, 1, 0]
    assert t[:, 9:].tolist() == [[0, 0, 1], [0, 0, 1]]  # Triangle::new normal


def test_png_writer_roundtrip(pt, tmp_path):
    import zlib
    buf = np.array([0x00FF0000, 0x0000FF00, 0x000000FF, 0x00B4B4B4, 0, 0x00FFFFFF], np.uint32)
    path = str(tmp_path / "o.png")
    pt.save_image(path, buf, 3, 2)
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(raw):
        n = int.from_bytes(raw[pos:pos + 4], "big")
        typ, data = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + data) == int.from_bytes(raw[pos + 8 + n:pos + 12 + n], "big")
        if typ == b"IHDR":
            ihdr = data
        if typ == b"IDAT":
            idat += data
        pos += 12 + n
    assert int.from_bytes(ihdr[:4], "big") == 3 and int.from_bytes(ihdr[4:8], "big") == 2 and ihdr[8:10] == b"\x08\x02"
    px = zlib.decompress(idat)
    assert px == bytes([0, 255, 0, 0, 0, 255, 0, 0, 0, 255, 0, 180, 180, 180, 0, 0, 0, 255, 255, 255])


def test_hdr_reader(pt, tmp_path):
    # flat (non-RLE) RGBE: (128,64,32,129) -> mantissa * 2^(129-136)
    p = tmp_path / "e.hdr"
    p.write_bytes(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 1 +X 2\n" + bytes([128, 64, 32, 129, 0, 0, 0, 0]))
    s = pt.Scene()
    s.set_sky_hdr_file(str(p))
    sky = s.sky
    assert sky.shape == (1, 2, 3)
    assert sky[0, 0].tolist() == [1.0, 0.5, 0.25] and sky[0, 1].tolist() == [0, 0, 0]


def test_synthetic_scene_shape(pt):
    s = pt.synthetic_scene(cells=20)
    assert s.settings == (3840, 2160, 256, 16)
    assert [o.type for o in s.objects] == [pt.OBJ_MESH, pt.OBJ_SPHERE, pt.OBJ_CUBE, pt.OBJ_QUAD]
    t = s.mesh(0)
    assert t.shape == (800, 12)
    assert t[:, [0, 3, 6]].min() == -50 and t[:, [0, 3, 6]].max() == 50


def _wo3_reference_misread(path):
    """Index triples exactly as Mesh::from_wo3 reads them (mesh_object.rs:188-190): 12 consecutive bytes per triangle
    although the records are 16 bytes long; triples with an index out of range are dropped (:201-215)."""
    raw = open(path, "rb").read()
    nv = int(np.frombuffer(raw, "<u8", 1, 0)[0])
    verts = np.frombuffer(raw, "<f4", nv * 8, 8).reshape(nv, 8)[:, :3]
    off = 8 + nv * 32
    nt = int(np.frombuffer(raw, "<u8", 1, off)[0])
    words = np.frombuffer(raw, "<u4", nt * 4, off + 8)
    wrong = words[:nt * 3].reshape(nt, 3)
    right = words.reshape(nt, 4)[:, :3]
    return verts, wrong[(wrong < nv).all(1)], right


def test_shipped_teapot_scene_strict_and_extended(pt, scenes_dir):
    path = os.path.join(scenes_dir, "teapot", "scene.json")
    # the reference cannot load its own teapot scene: `infinite_sphere` is not an ObjectConfigVariant (parser.rs:136-165)
    with pytest.raises(RuntimeError, match="unknown variant `infinite_sphere`"):
        pt.load_scene_from_json(path)
    s = pt.load_scene_from_json(path, pt.LOAD_INFINITE_SPHERE_SKY | pt.LOAD_WO3_STRIDE16)
    assert s.settings == (1280, 720, 64, 64)
    assert [o.type for o in s.objects] == [pt.OBJ_QUAD, pt.OBJ_MESH, pt.OBJ_MESH]
    assert s.sky is not None and s.sky.shape == (512, 1024, 3)
    mats = s.materials
    assert mats[s.objects[1].material].type == pt.MAT_PLASTIC and mats[s.objects[0].material].type == pt.MAT_LAMBERT_CHECKER
    # WO3: Tungsten's 16-byte records vs the reference's 12-byte mis-read, against an independent numpy reading
    for obj, name in [(s.objects[1], "Mesh001.wo3"), (s.objects[2], "Mesh000.wo3")]:
        verts, wrong, right = _wo3_reference_misread(os.path.join(scenes_dir, "teapot", "models", name))
        tris = s.mesh(obj.mesh)
        assert 0 < len(tris) <= len(right)
        # every loaded triangle is one of the file's (degenerate ones are filtered)
        v0 = {tuple(np.round(verts[t].ravel(), 6)) for t in right}
        assert all(tuple(np.round(tr[:9], 6)) in v0 for tr in tris[:200])
    strict_meshes = pt.load_scene_from_json(path, pt.LOAD_INFINITE_SPHERE_SKY)
    for obj, name in [(strict_meshes.objects[1], "Mesh001.wo3"), (strict_meshes.objects[2], "Mesh000.wo3")]:
        verts, wrong, right = _wo3_reference_misread(os.path.join(scenes_dir, "teapot", "models", name))
        tris = strict_meshes.mesh(obj.mesh)
        assert len(tris) <= len(wrong) < len(right)
        w0 = {tuple(np.round(verts[t].ravel(), 6)) for t in wrong}
        assert all(tuple(np.round(tr[:9], 6)) in w0 for tr in tris[:200])
    # unknown types: fatal by default, skipped on request
    with pytest.raises(RuntimeError, match="unknown variant"):
        pt.load_scene_from_json(path, pt.LOAD_WO3_STRIDE16)
    skipped = pt.load_scene_from_json(path, pt.LOAD_SKIP_UNKNOWN)
    assert [o.type for o in skipped.objects] == [pt.OBJ_QUAD, pt.OBJ_MESH, pt.OBJ_MESH] and skipped.sky is None


def test_reference_bvh_restatement_radix_equals_stable_sort(scenes_dir):
    """pt_build sorts big nodes with an LSD radix sort on the centroid's order-preserving integer image; it must produce
    the permutation std::stable_sort produces (dead mask, DFS leaf order, node / leaf / depth counts), here on the two
    WO3 meshes of the shipped teapot scene (76,968 and 47,872 triangles).  The switch is read once per process."""
    import subprocess
    import sys
    prog = (
        "import sys, hashlib; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import ptload; pt = ptload.load()\n"
        "from bindings import SimScene\n"
        "s = pt.load_scene_from_json(%r, pt.LOAD_INFINITE_SPHERE_SKY | pt.LOAD_WO3_STRIDE16)\n"
        "sim = SimScene(s)\n"
        "for k, o in enumerate(s.objects):\n"
        "    if o.type == pt.OBJ_MESH:\n"
        "        info, dead, order = sim.mesh_info(pt, k)\n"
        "        print(info.ref_nodes, info.ref_leaves, info.ref_depth, info.live_triangles, hashlib.md5(dead.tobytes()).hexdigest(),\n"
        "              hashlib.md5(order.tobytes()).hexdigest())\n"
    ) % (os.path.dirname(scenes_dir), os.path.join(os.path.dirname(scenes_dir), "tests"), os.path.join(scenes_dir, "teapot", "scene.json"))
    prog += (
        "import numpy as np\n"
        "rng = np.random.default_rng(5)\n"
        "o = rng.uniform(-30, 30, (20000, 3)).astype(np.float32); d = rng.normal(size=(20000, 3)).astype(np.float32)\n"
        "d /= np.linalg.norm(d, axis=1, keepdims=True)\n"
        "hits, st = sim.intersect(pt, o, d)\n"
        "print(hashlib.md5(hits.tobytes()).hexdigest(), int((hits['object'] >= 0).sum()))\n"
    )
    outs = []
    for var, val in (("", ""), ("PTC_REF_STABLE_SORT", "1"), ("PTC_COLLAPSE", "greedy")):
        env = dict(os.environ)
        env.pop("PTC_REF_STABLE_SORT", None)
        env.pop("PTC_COLLAPSE", None)
        if var:
            env[var] = val
        outs.append(subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, check=True).stdout)
    # same reference-BVH facts with either sort, and the same hit records whichever way the wide tree was collapsed
    # (the dynamic-programme collapse is the default, PTC_COLLAPSE=greedy the first version)
    assert outs[0] == outs[1] == outs[2] and len(outs[0].splitlines()) == 3
    assert int(outs[0].splitlines()[2].split()[1]) > 1000
